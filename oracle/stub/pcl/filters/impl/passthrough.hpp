// Stand-in for <pcl/filters/impl/passthrough.hpp>: the reference includes it but uses nothing from it on this path.  See oracle/stub/README.md.
#pragma once

// Stand-in for <pcl/io/io.h>: the reference includes it but uses nothing from it on this path.  See oracle/stub/README.md.
#pragma once

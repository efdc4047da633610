// Stand-in for <pcl/registration/icp_nl.h>: included by TopPartRegistration.cpp, nothing of it is used.  See ../../README.md.
#pragma once

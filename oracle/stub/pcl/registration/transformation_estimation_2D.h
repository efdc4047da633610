// Stand-in for <pcl/registration/transformation_estimation_2D.h>: included by TopPartRegistration.cpp, nothing of it is used.  See ../../README.md.
#pragma once

// Stand-in for <pcl/filters/statistical_outlier_removal.h>: included by TopPartRegistration.cpp, nothing of it is used.  See ../../README.md.
#pragma once

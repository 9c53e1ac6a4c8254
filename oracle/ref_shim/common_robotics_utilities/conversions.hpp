// TEST INFRASTRUCTURE ONLY -- empty stand-in for common_robotics_utilities/conversions.hpp: the
// reference's voxel_raycasting_test.cpp includes it and uses nothing from it.
#pragma once

// Empty stand-in so that the reference's Kinematics.{h,cpp} (which include ROS
// message headers they never use) compile headless.  Test infrastructure only.
#pragma once

// TEST INFRASTRUCTURE stub: only ros::Time / ros::Duration are needed by the trajectory code
#pragma once
#include "time.h"

// TEST INFRASTRUCTURE stub
#pragma once
#include <vector>
#include "Event.h"
namespace dvs_msgs { struct EventArray { std::vector<Event> events; typedef const EventArray* ConstPtr; }; }

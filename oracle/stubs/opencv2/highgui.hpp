// TEST INFRASTRUCTURE stub: see core.hpp
#pragma once
#include "core.hpp"

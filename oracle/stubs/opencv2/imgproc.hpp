// TEST INFRASTRUCTURE stub: see core.hpp
#pragma once
#include "core.hpp"
#include "ops_stub.hpp"

// TEST INFRASTRUCTURE stub
#pragma once

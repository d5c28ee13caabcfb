// TEST INFRASTRUCTURE stub
#pragma once
namespace image_geometry { class PinholeCameraModel; }

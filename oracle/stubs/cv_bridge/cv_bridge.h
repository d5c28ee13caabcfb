// TEST INFRASTRUCTURE stub
#pragma once
namespace cv_bridge { struct CvImage {}; }

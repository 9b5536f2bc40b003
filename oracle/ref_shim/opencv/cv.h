// TEST INFRASTRUCTURE ONLY: stand-in for the two OpenCV names src/mapFeatures.cpp touches (descriptor bookkeeping that the LBA
// path never reaches): an opaque cv::Mat and cv::norm(..., NORM_HAMMING).  Written from scratch; see ../standin.h.
#pragma once
namespace cv {
class Mat {};
enum { NORM_HAMMING = 6 };
inline double norm(const Mat &, const Mat &, int) { return 0.0; }
}

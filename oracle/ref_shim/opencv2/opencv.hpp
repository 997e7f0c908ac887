// STAND-IN for <opencv2/opencv.hpp> — TEST INFRASTRUCTURE ONLY (see ../NumTypes.h). src/internal/FrameHessian.cc touches OpenCV in one
// debug branch of makeImages (a display copy of the image, taken only when setting_showLoopClosing is on); this is the type it names.
#pragma once
#include <vector>
typedef unsigned char uchar;
#define CV_8UC3 16
namespace cv {
struct Mat {
    std::vector<uchar> buf; uchar *data = nullptr; int rows = 0, cols = 0;
    Mat() {}
    Mat(int r, int c, int) : buf((size_t) r * c * 3), rows(r), cols(c) { data = buf.data(); }
    Mat(const Mat &o) : buf(o.buf), rows(o.rows), cols(o.cols) { data = buf.data(); }
    Mat &operator=(const Mat &o) { buf = o.buf; rows = o.rows; cols = o.cols; data = buf.data(); return *this; }
};
}

// TEST-ONLY functional stand-in (see core.hpp): the one imgproc function the adapter's SIMPLE detector uses.
#pragma once
#include "core.hpp"

namespace cv
{
    // implemented over the C oracle in oracle/ref_glue/opencv_over_oracle.cpp (the adapter does not call it)
    void cornerSubPix(InputArray image, std::vector<Point2f>& corners, Size winSize, Size zeroZone, TermCriteria criteria);

    // filled disc (thickness < 0) on a CV_8UC1 image, clipped: pixels with dx^2 + dy^2 <= r^2 around the centre (a cv::Point: the caller's Point2f
    // converts with cvRound) -- the set OpenCV's filled midpoint circle covers for the radii zenslam uses (cell / 2 = 8 .. 32);
    // tests/test_gpu_adapter.py compares the resulting detections with the python mirror, whose mask is pinned to cv2.circle
    inline void circle(Mat& img, Point center, int radius, const Scalar& color, int thickness = 1)
    {
        CV_Assert(thickness < 0 && img.type() == CV_8UC1);
        const int  cx = center.x, cy = center.y;
        const auto v  = static_cast<uchar>(color[0]);
        for (int y = std::max(0, cy - radius); y <= std::min(img.rows - 1, cy + radius); ++y)
            for (int x = std::max(0, cx - radius); x <= std::min(img.cols - 1, cx + radius); ++x)
                if ((x - cx) * (x - cx) + (y - cy) * (y - cy) <= radius * radius)
                    img.at<uchar>(y, x) = v;
    }
}

// TEST-ONLY stand-in (see core.hpp): the two imgproc declarations the adapter's SIMPLE detector uses.
#pragma once
#include "core.hpp"

namespace cv
{
    void circle(Mat& img, Point2f center, int radius, const Scalar& color, int thickness = 1);
}

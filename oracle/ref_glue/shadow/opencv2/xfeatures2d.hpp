// TEST INFRASTRUCTURE (oracle/ref_glue): cv::xfeatures2d::FREAK is only named by the detector constructors (descriptor: FREAK
// is outside the CUDA path); creating one here is an error
#pragma once
#include <opencv2/features2d.hpp>
namespace cv::xfeatures2d
{
    class FREAK : public Feature2D
    {
    public:
        static Ptr<FREAK> create() { CV_Error(Error::StsNotImplemented, "FREAK is not part of the oracle"); }
    };
}

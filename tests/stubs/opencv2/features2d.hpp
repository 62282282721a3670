// TEST-ONLY stand-in, see core.hpp
#pragma once
#include "core.hpp"

namespace cv
{
    class DescriptorMatcher
    {
    public:
        virtual ~DescriptorMatcher() = default;
        [[nodiscard]] virtual bool isMaskSupported() const = 0;
        [[nodiscard]] virtual Ptr<DescriptorMatcher> clone(bool emptyTrainData = false) const = 0;
    protected:
        virtual void knnMatchImpl(InputArray queryDescriptors, std::vector<std::vector<DMatch>>& matches, int k,
                                  InputArrayOfArrays masks = _InputArray(), bool compactResult = false) = 0;
        virtual void radiusMatchImpl(InputArray queryDescriptors, std::vector<std::vector<DMatch>>& matches, float maxDistance,
                                     InputArrayOfArrays masks = _InputArray(), bool compactResult = false) = 0;
        std::vector<Mat> trainDescCollection;
    };
}

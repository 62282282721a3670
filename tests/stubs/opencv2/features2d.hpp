// TEST-ONLY functional stand-in (see core.hpp): cv::DescriptorMatcher with the public, non-virtual front of OpenCV 4.x --
// add / clear / empty / train, knnMatch and match in both their (query, train, ...) and (query, ...) forms -- forwarding to
// the virtual knnMatchImpl the way the OpenCV documentation describes: the two-image forms clone(true) the matcher, add the
// train descriptors and call the one-image form with std::vector<Mat>(1, mask); match() is knnMatch(k = 1, compactResult =
// true) flattened.  zenslam::matcher calls exactly these (matcher.cpp:65,79), so a subclass that mishandles the masks vector
// fails here like it would against the real library.
#pragma once
#include "core.hpp"

namespace cv
{
    class DescriptorMatcher
    {
    public:
        virtual ~DescriptorMatcher() = default;

        virtual void add(InputArrayOfArrays descriptors)
        {
            std::vector<Mat> v;
            descriptors.getMatVector(v);
            trainDescCollection.insert(trainDescCollection.end(), v.begin(), v.end());
        }
        [[nodiscard]] const std::vector<Mat>& getTrainDescriptors() const { return trainDescCollection; }
        virtual void clear() { trainDescCollection.clear(); }
        [[nodiscard]] virtual bool empty() const { return trainDescCollection.empty(); }
        [[nodiscard]] virtual bool isMaskSupported() const = 0;
        virtual void train() { }

        void knnMatch(InputArray query, InputArray train_descriptors, std::vector<std::vector<DMatch>>& matches, int k,
                      InputArray mask = noArray(), bool compactResult = false) const
        {
            Ptr<DescriptorMatcher> temp = clone(true);
            temp->add(std::vector<Mat>(1, train_descriptors.getMat()));
            temp->knnMatch(query, matches, k, std::vector<Mat>(1, mask.getMat()), compactResult);
        }

        void knnMatch(InputArray query, std::vector<std::vector<DMatch>>& matches, int k, InputArrayOfArrays masks = noArray(),
                      bool compactResult = false)
        {
            if (empty() || query.empty())
                return;
            CV_Assert(k > 0);
            check_masks(masks, query.size().height);
            train();
            knnMatchImpl(query, matches, k, masks, compactResult);
        }

        void match(InputArray query, InputArray train_descriptors, std::vector<DMatch>& matches, InputArray mask = noArray()) const
        {
            Ptr<DescriptorMatcher> temp = clone(true);
            temp->add(std::vector<Mat>(1, train_descriptors.getMat()));
            temp->match(query, matches, std::vector<Mat>(1, mask.getMat()));
        }

        void match(InputArray query, std::vector<DMatch>& matches, InputArrayOfArrays masks = noArray())
        {
            std::vector<std::vector<DMatch>> knn;
            knnMatch(query, knn, 1, masks, true);
            matches.clear();
            for (const auto& row : knn)
                matches.insert(matches.end(), row.begin(), row.end());
        }

        [[nodiscard]] virtual Ptr<DescriptorMatcher> clone(bool emptyTrainData = false) const = 0;

    protected:
        virtual void knnMatchImpl(InputArray queryDescriptors, std::vector<std::vector<DMatch>>& matches, int k,
                                  InputArrayOfArrays masks = noArray(), bool compactResult = false) = 0;
        virtual void radiusMatchImpl(InputArray queryDescriptors, std::vector<std::vector<DMatch>>& matches, float maxDistance,
                                     InputArrayOfArrays masks = noArray(), bool compactResult = false) = 0;

        // OpenCV only validates the masks of a matcher that supports them
        void check_masks(InputArrayOfArrays masks, int query_count) const
        {
            if (!isMaskSupported() || masks.empty())
                return;
            std::vector<Mat> v;
            masks.getMatVector(v);
            CV_Assert(v.size() == trainDescCollection.size());
            for (size_t i = 0; i < v.size(); ++i)
                CV_Assert(v[i].empty() || (v[i].type() == CV_8UC1 && v[i].rows == query_count && v[i].cols == trainDescCollection[i].rows));
        }

        std::vector<Mat> trainDescCollection;
    };
}

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
    // cv::Feature2D and the detectors / describers the reference constructs (keypoint_detector_grid.cpp:12-36).  Declarations
    // only: the adapter never calls them; oracle/ref_glue/opencv_over_oracle.cpp implements them over the C oracle so that the
    // reference's own detector sources can be compiled and run (oracle/_ref).
    class Feature2D
    {
    public:
        virtual ~Feature2D() = default;
        virtual void detect(InputArray image, std::vector<KeyPoint>& keypoints, InputArray mask = noArray());
        virtual void compute(InputArray image, std::vector<KeyPoint>& keypoints, OutputArray descriptors);
    };

    class FastFeatureDetector : public Feature2D
    {
    public:
        static Ptr<FastFeatureDetector> create(int threshold = 10, bool nonmaxSuppression = true);
        void detect(InputArray image, std::vector<KeyPoint>& keypoints, InputArray mask = noArray()) override;
    private:
        int _threshold { 10 };
    };

    class ORB : public Feature2D
    {
    public:
        enum ScoreType { HARRIS_SCORE = 0, FAST_SCORE = 1 };
        static Ptr<ORB> create(int nfeatures = 500, float scaleFactor = 1.2f, int nlevels = 8, int edgeThreshold = 31, int firstLevel = 0,
                               int WTA_K = 2, ScoreType scoreType = HARRIS_SCORE, int patchSize = 31, int fastThreshold = 20);
        void detect(InputArray image, std::vector<KeyPoint>& keypoints, InputArray mask = noArray()) override;
        void compute(InputArray image, std::vector<KeyPoint>& keypoints, OutputArray descriptors) override;
    private:
        int   _nfeatures { 500 }, _nlevels { 8 }, _edge { 31 }, _patch { 31 }, _fast_threshold { 20 };
        float _scale { 1.2f };
    };

    class SIFT : public Feature2D
    {
    public:
        static Ptr<SIFT> create();
    };

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

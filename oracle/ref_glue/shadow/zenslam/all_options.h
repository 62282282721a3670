// TEST INFRASTRUCTURE (oracle/ref_glue): stands in for the reference's zenslam/all_options.h while the reference's OWN detector
// sources are compiled for oracle/_ref.  The real header pulls yaml-cpp, the IMU integrator and the I/O layer in; the detector
// classes only use detection_options, which comes from the reference's own header below.
#pragma once
#include "zenslam/detection/detection_options.h"
#include "zenslam/tracking_options.h"

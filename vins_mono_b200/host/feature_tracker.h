// Host shim with the reference's class surface: drop-in for feature_tracker/src/feature_tracker.h:28-64.
// feature_tracker_node.cpp compiles against this class unchanged apart from the include: readImage() takes any
// cv::Mat-like object exposing .data / .rows / .cols / .step (8UC1); the public result vectors keep their names.
// The work happens behind the vt_* C ABI (include/vinsb200/tracker.h) on the GPU; there is no CPU path.
#pragma once
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "vinsb200/tracker.h"

namespace vinsb200 {

struct Point2f {  // layout-compatible with cv::Point2f
    float x, y;
};

extern bool PUB_THIS_FRAME;  // the reference's global (feature_tracker/src/parameters.cpp:19)

class FeatureTracker {
  public:
    FeatureTracker() = default;
    ~FeatureTracker() { vt_destroy(h_); }
    FeatureTracker(const FeatureTracker&) = delete;
    FeatureTracker& operator=(const FeatureTracker&) = delete;

    // readParameters + readIntrinsicParameter: the YAML values, already parsed by the node
    void configure(const vt_config& cfg) {
        vt_destroy(h_);
        h_ = nullptr;
        if (vt_create(&cfg, &h_) != VT_OK) throw std::runtime_error("vt_create failed (no CUDA device?)");
        cap_ = cfg.max_cnt;
    }

    template <class MatLike>
    void readImage(const MatLike& _img, double _cur_time) {
        check(vt_read_image(h_, _img.data, (size_t)_img.step, _cur_time, PUB_THIS_FRAME ? 1 : 0));
        cur_time = _cur_time;
        fetch();
    }

    // ids are assigned inside readImage in the reference's order; kept for source compatibility with the node's loop
    bool updateID(unsigned int i) { return i < ids.size(); }

    std::vector<Point2f> cur_pts, cur_un_pts, pts_velocity;
    std::vector<int> ids, track_cnt;
    double cur_time = 0;

  private:
    void check(int rc) {
        if (rc < 0) throw std::runtime_error(std::string("vinsb200: ") + vt_last_error(h_));
    }
    void fetch() {
        const int n = vt_count(h_);
        cur_pts.resize(n); cur_un_pts.resize(n); pts_velocity.resize(n); ids.resize(n); track_cnt.resize(n);
        vt_get(h_, ids.data(), track_cnt.data(), &cur_pts[0].x, &cur_un_pts[0].x, &pts_velocity[0].x);
    }
    vt_tracker* h_ = nullptr;
    int cap_ = 0;
};

}  // namespace vinsb200

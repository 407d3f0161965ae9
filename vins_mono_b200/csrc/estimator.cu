// ve_* C ABI: host shim of the reference's Estimator / FeatureManager (vins_estimator/src/estimator.cpp,
// feature_manager.cpp) around the CUDA BA path.  Bookkeeping that the reference does on the host stays on the
// host (window arrays, feature tracks, slide/re-anchor logic, gauge re-anchoring in double2vector); IMU
// pre-integration, factor linearisation, the trust-region solve and the marginalisation run on the GPU with all
// problem data resident in HBM.  Per frame the PCIe traffic is the packed observation table + states down and
// the states back.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "ba_kernels.h"
#include "host_math.h"
#include "vinsb200/estimator.h"

using hm::Mat3;
using hm::Quat;
using hm::Vec3;

namespace {

struct FeaturePerFrame {  // feature_manager.h:19-46
    Vec3 point;
    double u, v, vx, vy, cur_td;
};
struct FeaturePerId {  // feature_manager.h:48-72
    int feature_id, start_frame;
    std::vector<FeaturePerFrame> feature_per_frame;
    int used_num = 0;
    double estimated_depth = -1.0;
    int solve_flag = 0;
    int endFrame() const { return start_frame + (int)feature_per_frame.size() - 1; }
};

struct SeedRow {
    double t;
    Vec3 P, V;
    Mat3 R;
};

struct PriorBlock {
    int type, index, off, size;  // type 0 pose, 1 speed-bias, 2 ex pose, 3 td; size = local size
};

template <class T>
struct DeviceBuf {
    T* p = nullptr;
    size_t n = 0;
    cudaError_t alloc(size_t count) {
        n = count;
        return cudaMalloc(&p, sizeof(T) * std::max<size_t>(count, 1));
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
    }
};

}  // namespace

struct ve_estimator {
    ve_config cfg{};
    std::string err;
    int W = 10;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[8] = {};  // 0-2 solve phases, 3 marginalisation done, 5 marginalisation start, 6 state upload of marginalize()
    // ---- Estimator state (estimator.h:65-115)
    int solver_flag = 0;           // INITIAL = 0, NON_LINEAR = 1
    int marginalization_flag = 0;  // MARGIN_OLD = 0, MARGIN_SECOND_NEW = 1
    Vec3 g;
    Mat3 ric;
    Vec3 tic;
    std::vector<Vec3> Ps, Vs, Bas, Bgs;
    std::vector<Mat3> Rs;
    double td = 0;
    Mat3 back_R0, last_R, last_R0;
    Vec3 back_P0, last_P, last_P0;
    std::vector<double> Headers;
    Vec3 acc_0, gyr_0;
    std::vector<std::vector<double>> dt_buf;
    std::vector<std::vector<Vec3>> acc_buf, gyr_buf;
    int frame_count = 0;
    bool first_imu = false, failure_occur = false;
    std::vector<FeaturePerId> feature;  // FeatureManager::feature (std::list in the reference; order preserved)
    int last_track_num = 0;
    std::vector<SeedRow> seeds;
    Vec3 seed_ba, seed_bg;
    // pre-integration slots: frame -> slot, per-slot host mirror of what the device slot was created with
    std::vector<int> slot_of;
    std::vector<bool> slot_valid;
    std::vector<size_t> flushed;    // samples of dt_buf[frame] already integrated on the device
    std::vector<double> sum_dt;     // per frame
    std::vector<Vec3> lin_acc, lin_gyr;  // linearized_acc / linearized_gyr per frame
    std::vector<bool> sqrt_dirty;   // per slot
    // prior (MarginalizationInfo) bookkeeping
    bool has_prior = false;
    int prior_n = 0, prior_buf = 0;
    std::vector<PriorBlock> prior_blocks;
    // summary
    int n_solves = 0, n_reboots = 0, last_landmarks = 0, last_visual = 0;
    vb::SolverState last_state{};
    float last_ms[4] = {0, 0, 0, 0};
    int last_launches = 0;
    vb::KernelProfile prof;
    size_t h2d_bytes = 0, d2h_bytes = 0;  // of the last process_image
    // ---- device memory
    int Lmax = 0, Mmax = 0, D = 0, nmax = 0;
    DeviceBuf<vb::PreInt> d_preint;
    DeviceBuf<double> d_samples;
    DeviceBuf<int> d_which;
    DeviceBuf<double> d_states[2];   // pose | sb | ex | td | lam
    DeviceBuf<double> d_acc[2];      // Hpp | gp | Hpl | Hll | gl | cost
    DeviceBuf<int> d_ints;           // lm_anchor | lm_start | ob_frame | imu_slot
    DeviceBuf<double> d_obs;         // lm_pts | lm_vel | lm_td | lm_row | ob_pts | ob_vel | ob_td | ob_row
    DeviceBuf<double> d_S, d_Spk, d_Hfull, d_gred, d_vec, d_work;
    DeviceBuf<vb::SolverState> d_st;
    DeviceBuf<double> d_prior[2];    // A | g0 | c0 | x0
    DeviceBuf<int> d_prior_i[2];     // type | index | off
    DeviceBuf<double> d_marg;        // Am | bm | Araw | graw
    DeviceBuf<int> d_marg_i;         // lms | col_lm
    // pinned staging
    double* h_states = nullptr;
    double* h_obs = nullptr;
    int* h_ints = nullptr;
    double* h_samples = nullptr;
    vb::PreInt* h_preint = nullptr;
    vb::SolverState* h_st = nullptr;
    double* h_prior = nullptr;
    int* h_prior_i = nullptr;
    int* h_marg_i = nullptr;
    int* h_which = nullptr;          // pinned staging of refresh_sqrt_info (h_marg_i may still be in flight)
    double* h_marg_out = nullptr;
    std::vector<double> prior_raw_A, prior_raw_b;  // last Schur complement before the eps floor
    double marg_sweeps[7] = {0, 0, 0, 0, 0, 0, 0};
    int sample_seg = 0, flushes_in_flight = 0;
    bool marg_pending = false;  // marginalisation kernels enqueued, results not yet read back
    int marg_n = 0;
    bool states_upload_pending = false;
};

namespace {

#define VE_CUDA(call)                                                        \
    do {                                                                     \
        cudaError_t e_ = (call);                                             \
        if (e_ != cudaSuccess) {                                             \
            e->err = std::string(#call) + ": " + cudaGetErrorString(e_);     \
            return VE_ERR_CUDA;                                              \
        }                                                                    \
    } while (0)

size_t states_doubles(const ve_estimator* e) { return (size_t)(e->W + 1) * 16 + 8 + e->Lmax; }
size_t acc_doubles(const ve_estimator* e) { return (size_t)e->D * e->D + e->D + (size_t)e->Lmax * e->D + 2 * (size_t)e->Lmax + 1; }

vb::BaStates states_view(const ve_estimator* e, int b) {
    double* p = e->d_states[b].p;
    const int F = e->W + 1;
    vb::BaStates s;
    s.pose = p;
    s.sb = p + 7 * F;
    s.ex = p + 16 * F;
    s.td = p + 16 * F + 7;
    s.lam = p + 16 * F + 8;
    return s;
}
bool usable(const ve_estimator* e, FeaturePerId& it) {  // the filter repeated all over feature_manager.cpp
    it.used_num = (int)it.feature_per_frame.size();
    return it.used_num >= 2 && it.start_frame < e->W - 2;
}

// ---- pre-integration slots ---------------------------------------------------------------------
int init_slot(ve_estimator* e, int frame, const Vec3& a0, const Vec3& g0, const Vec3& ba, const Vec3& bg) {
    const double A0[3] = {a0.x, a0.y, a0.z}, G0[3] = {g0.x, g0.y, g0.z}, BA[3] = {ba.x, ba.y, ba.z}, BG[3] = {bg.x, bg.y, bg.z};
    const int slot = e->slot_of[frame];
    vb::launch_preint_init(e->d_preint.p + slot, A0, G0, BA, BG, e->stream);  // values travel as kernel arguments
    e->last_launches++;
    e->slot_valid[frame] = true;
    e->flushed[frame] = 0;
    e->sum_dt[frame] = 0;
    e->lin_acc[frame] = a0;
    e->lin_gyr[frame] = g0;
    e->sqrt_dirty[slot] = true;
    return VE_OK;
}

// Integrates the not-yet-integrated samples of `frame`'s buffers into its device slot.
int flush_frame(ve_estimator* e, int frame) {
    const size_t n = e->dt_buf[frame].size();
    if (!e->slot_valid[frame] || e->flushed[frame] >= n) return VE_OK;
    const size_t k0 = e->flushed[frame], cnt = n - k0;
    if (cnt > 512) {
        e->err = "too many IMU samples in one interval";
        return VE_ERR_CAPACITY;
    }
    // ring of 8 staging segments: no host wait here; a segment is reused 8 flushes later, by which time at least one
    // per-frame synchronisation (the solve's state read-back) has drained the stream
    if (++e->flushes_in_flight > 8) {  // every staging segment may still be waiting for its copy
        VE_CUDA(cudaStreamSynchronize(e->stream));
        e->flushes_in_flight = 1;
    }
    const int seg = e->sample_seg;
    e->sample_seg = (e->sample_seg + 1) & 7;
    double* hs = e->h_samples + (size_t)seg * 7 * 512;
    double* ds = e->d_samples.p + (size_t)seg * 7 * 512;
    for (size_t k = 0; k < cnt; k++) {
        double* s = hs + 7 * k;
        s[0] = e->dt_buf[frame][k0 + k];
        const Vec3 &a = e->acc_buf[frame][k0 + k], &w = e->gyr_buf[frame][k0 + k];
        s[1] = a.x; s[2] = a.y; s[3] = a.z; s[4] = w.x; s[5] = w.y; s[6] = w.z;
    }
    VE_CUDA(cudaMemcpyAsync(ds, hs, sizeof(double) * 7 * cnt, cudaMemcpyHostToDevice, e->stream));
    e->prof.begin(e->stream);
    vb::launch_preint_push(e->d_preint.p + e->slot_of[frame], (int)cnt, ds, e->cfg.acc_n, e->cfg.gyr_n, e->cfg.acc_w,
                           e->cfg.gyr_w, e->stream);
    e->prof.end(6, e->stream);
    e->h2d_bytes += sizeof(double) * 7 * cnt;
    e->last_launches++;
    e->flushed[frame] = n;
    e->sqrt_dirty[e->slot_of[frame]] = true;
    return VE_OK;
}

int refresh_sqrt_info(ve_estimator* e) {
    int cnt = 0;
    for (int f = 1; f <= e->W; f++) {
        const int s = e->slot_of[f];
        if (e->slot_valid[f] && e->sqrt_dirty[s] && e->dt_buf[f].size() > 0) {
            e->h_which[cnt++] = s;
            e->sqrt_dirty[s] = false;
        }
    }
    if (!cnt) return VE_OK;
    // no host wait: h_which is rewritten only after the next solve's read-back has drained the stream
    VE_CUDA(cudaMemcpyAsync(e->d_which.p, e->h_which, sizeof(int) * cnt, cudaMemcpyHostToDevice, e->stream));
    e->prof.begin(e->stream);
    vb::launch_sqrt_info(e->d_preint.p, e->d_which.p, cnt, e->stream);
    e->prof.end(7, e->stream);
    e->last_launches++;
    return VE_OK;
}

// ---- FeatureManager (feature_manager.cpp) ------------------------------------------------------
double compensated_parallax2(const FeaturePerId& it, int frame_count) {
    const FeaturePerFrame& fi = it.feature_per_frame[frame_count - 2 - it.start_frame];
    const FeaturePerFrame& fj = it.feature_per_frame[frame_count - 1 - it.start_frame];
    const double dep_i = fi.point.z;
    const double du = fi.point.x / dep_i - fj.point.x, dv = fi.point.y / dep_i - fj.point.y;
    return std::max(0.0, std::sqrt(std::min(du * du + dv * dv, du * du + dv * dv)));
}

bool add_feature_check_parallax(ve_estimator* e, int n, const int* order, const int* ids, const double* d7, double td) {
    double parallax_sum = 0;
    int parallax_num = 0;
    e->last_track_num = 0;
    for (int k = 0; k < n; k++) {  // ascending feature id (std::map order)
        const int i = order[k];
        const double* d = d7 + 7 * i;
        const FeaturePerFrame f{Vec3(d[0], d[1], d[2]), d[3], d[4], d[5], d[6], td};
        auto it = std::find_if(e->feature.begin(), e->feature.end(), [&](const FeaturePerId& x) { return x.feature_id == ids[i]; });
        if (it == e->feature.end()) {
            FeaturePerId nf;
            nf.feature_id = ids[i];
            nf.start_frame = e->frame_count;
            nf.feature_per_frame.push_back(f);
            e->feature.push_back(nf);
        } else {
            it->feature_per_frame.push_back(f);
            e->last_track_num++;
        }
    }
    const int fc = e->frame_count;
    if (fc < 2 || e->last_track_num < 20) return true;
    for (auto& it : e->feature)
        if (it.start_frame <= fc - 2 && it.start_frame + (int)it.feature_per_frame.size() - 1 >= fc - 1) {
            parallax_sum += compensated_parallax2(it, fc);
            parallax_num++;
        }
    if (parallax_num == 0) return true;
    return parallax_sum / parallax_num >= e->cfg.keyframe_parallax / e->cfg.focal_length;
}

void triangulate(ve_estimator* e) {
    for (auto& it : e->feature) {
        if (!usable(e, it)) continue;
        if (it.estimated_depth > 0) continue;
        const int imu_i = it.start_frame;
        int imu_j = imu_i - 1;
        double A[64 * 4];
        int row = 0;
        const Vec3 t0 = e->Ps[imu_i] + e->Rs[imu_i] * e->tic;
        const Mat3 R0 = e->Rs[imu_i] * e->ric;
        for (auto& f : it.feature_per_frame) {
            imu_j++;
            const Vec3 t1 = e->Ps[imu_j] + e->Rs[imu_j] * e->tic;
            const Mat3 R1 = e->Rs[imu_j] * e->ric;
            const Vec3 t = R0.T() * (t1 - t0);
            const Mat3 Rt = (R0.T() * R1).T();
            const Vec3 mt = (Rt * t) * -1.0;
            double P[3][4];
            for (int i = 0; i < 3; i++) {
                for (int j = 0; j < 3; j++) P[i][j] = Rt(i, j);
                P[i][3] = mt[i];
            }
            const double nn = f.point.norm();
            const Vec3 fn = f.point * (1.0 / nn);
            for (int c = 0; c < 4; c++) {
                A[4 * row + c] = fn.x * P[2][c] - fn.z * P[0][c];
                A[4 * (row + 1) + c] = fn.y * P[2][c] - fn.z * P[1][c];
            }
            row += 2;
        }
        double v[4];
        hm::null_direction4(A, row, v);
        it.estimated_depth = v[2] / v[3];
        if (it.estimated_depth < 0.1) it.estimated_depth = e->cfg.init_depth;
    }
}

void remove_back_shift_depth(ve_estimator* e, const Mat3& marg_R, const Vec3& marg_P, const Mat3& new_R, const Vec3& new_P) {
    std::vector<FeaturePerId> keep;
    keep.reserve(e->feature.size());
    for (auto& it : e->feature) {
        if (it.start_frame != 0) {
            it.start_frame--;
            keep.push_back(std::move(it));
            continue;
        }
        const Vec3 uv_i = it.feature_per_frame[0].point;
        it.feature_per_frame.erase(it.feature_per_frame.begin());
        if (it.feature_per_frame.size() < 2) continue;
        const Vec3 pts_i = uv_i * it.estimated_depth;
        const Vec3 w_pts_i = marg_R * pts_i + marg_P;
        const Vec3 pts_j = new_R.T() * (w_pts_i - new_P);
        it.estimated_depth = pts_j.z > 0 ? pts_j.z : e->cfg.init_depth;
        keep.push_back(std::move(it));
    }
    e->feature.swap(keep);
}
void remove_back(ve_estimator* e) {
    std::vector<FeaturePerId> keep;
    for (auto& it : e->feature) {
        if (it.start_frame != 0)
            it.start_frame--;
        else {
            it.feature_per_frame.erase(it.feature_per_frame.begin());
            if (it.feature_per_frame.empty()) continue;
        }
        keep.push_back(std::move(it));
    }
    e->feature.swap(keep);
}
void remove_front(ve_estimator* e, int frame_count) {
    std::vector<FeaturePerId> keep;
    for (auto& it : e->feature) {
        if (it.start_frame == frame_count)
            it.start_frame--;
        else if (it.endFrame() >= frame_count - 1) {
            const int j = e->W - 1 - it.start_frame;
            it.feature_per_frame.erase(it.feature_per_frame.begin() + j);
            if (it.feature_per_frame.empty()) continue;
        }
        keep.push_back(std::move(it));
    }
    e->feature.swap(keep);
}

// ---- Estimator ---------------------------------------------------------------------------------
void set_parameter(ve_estimator* e) {
    e->tic = Vec3(e->cfg.tic[0], e->cfg.tic[1], e->cfg.tic[2]);
    std::memcpy(e->ric.m, e->cfg.ric, sizeof(e->ric.m));
    e->td = e->cfg.td;
    e->g = Vec3(0, 0, e->cfg.g_norm);
}

void clear_state(ve_estimator* e) {
    if (e->stream) cudaStreamSynchronize(e->stream);  // nothing of the old state may still be in flight
    e->marg_pending = false;
    e->states_upload_pending = false;
    e->flushes_in_flight = 0;
    for (int i = 0; i <= e->W; i++) {
        e->Rs[i] = Mat3();
        e->Ps[i] = e->Vs[i] = e->Bas[i] = e->Bgs[i] = Vec3();
        e->dt_buf[i].clear();
        e->acc_buf[i].clear();
        e->gyr_buf[i].clear();
        e->slot_valid[i] = false;
        e->slot_of[i] = i;
        e->flushed[i] = 0;
        e->sum_dt[i] = 0;
    }
    e->tic = Vec3();
    e->ric = Mat3();
    e->solver_flag = 0;
    e->first_imu = false;
    e->frame_count = 0;
    e->td = e->cfg.td;
    e->has_prior = false;
    e->prior_n = 0;
    e->prior_blocks.clear();
    e->feature.clear();
    e->failure_occur = false;
}

int process_imu(ve_estimator* e, double dt, const Vec3& acc, const Vec3& gyr) {
    if (!e->first_imu) {
        e->first_imu = true;
        e->acc_0 = acc;
        e->gyr_0 = gyr;
    }
    const int j = e->frame_count;
    if (!e->slot_valid[j]) {
        const int rc = init_slot(e, j, e->acc_0, e->gyr_0, e->Bas[j], e->Bgs[j]);
        if (rc) return rc;
    }
    if (j != 0) {
        e->dt_buf[j].push_back(dt);
        e->acc_buf[j].push_back(acc);
        e->gyr_buf[j].push_back(gyr);
        e->sum_dt[j] += dt;
        const Vec3 un_acc_0 = e->Rs[j] * (e->acc_0 - e->Bas[j]) - e->g;
        const Vec3 un_gyr = 0.5 * (e->gyr_0 + gyr) - e->Bgs[j];
        e->Rs[j] = e->Rs[j] * hm::deltaQ_R(un_gyr * dt);
        const Vec3 un_acc_1 = e->Rs[j] * (acc - e->Bas[j]) - e->g;
        const Vec3 un_acc = 0.5 * (un_acc_0 + un_acc_1);
        e->Ps[j] += dt * e->Vs[j] + 0.5 * dt * dt * un_acc;
        e->Vs[j] += dt * un_acc;
    }
    e->acc_0 = acc;
    e->gyr_0 = gyr;
    return VE_OK;
}

int initial_from_seed(ve_estimator* e, bool* ok) {
    *ok = false;
    for (int i = 0; i <= e->W; i++) {
        const SeedRow* s = nullptr;
        for (auto& c : e->seeds)
            if (std::fabs(c.t - e->Headers[i]) < 1e-6) s = &c;
        if (!s) return VE_OK;
        e->Ps[i] = s->P; e->Rs[i] = s->R; e->Vs[i] = s->V; e->Bas[i] = e->seed_ba; e->Bgs[i] = e->seed_bg;
    }
    for (int i = 0; i <= e->W; i++) {  // IntegrationBase::repropagate(Bas[i], Bgs[i])
        if (!e->slot_valid[i]) continue;
        const Vec3 la = e->lin_acc[i], lg = e->lin_gyr[i];
        int rc = init_slot(e, i, la, lg, e->Bas[i], e->Bgs[i]);
        if (rc) return rc;
        e->sum_dt[i] = 0;
        for (double v : e->dt_buf[i]) e->sum_dt[i] += v;
        rc = flush_frame(e, i);
        if (rc) return rc;
    }
    for (auto& it : e->feature) it.estimated_depth = -1;
    triangulate(e);
    *ok = true;
    return VE_OK;
}

bool failure_detection(ve_estimator* e) {
    const int W = e->W;
    if (e->Bas[W].norm() > 2.5) return true;
    if (e->Bgs[W].norm() > 1.0) return true;
    const Vec3 tmp_P = e->Ps[W];
    if ((tmp_P - e->last_P).norm() > 5) return true;
    if (std::fabs(tmp_P.z - e->last_P.z) > 1) return true;
    return false;
}

// vector2double (estimator.cpp:486-528): packs the window into the staging layout pose|sb|ex|td|lam
void pack_states(ve_estimator* e, double* out, int* n_lam) {
    const int F = e->W + 1;
    for (int i = 0; i < F; i++) {
        const Quat q = Quat::FromR(e->Rs[i]);
        double* p = out + 7 * i;
        p[0] = e->Ps[i].x; p[1] = e->Ps[i].y; p[2] = e->Ps[i].z; p[3] = q.x; p[4] = q.y; p[5] = q.z; p[6] = q.w;
        double* s = out + 7 * F + 9 * i;
        s[0] = e->Vs[i].x; s[1] = e->Vs[i].y; s[2] = e->Vs[i].z;
        s[3] = e->Bas[i].x; s[4] = e->Bas[i].y; s[5] = e->Bas[i].z;
        s[6] = e->Bgs[i].x; s[7] = e->Bgs[i].y; s[8] = e->Bgs[i].z;
    }
    const Quat q = Quat::FromR(e->ric);
    double* x = out + 16 * F;
    x[0] = e->tic.x; x[1] = e->tic.y; x[2] = e->tic.z; x[3] = q.x; x[4] = q.y; x[5] = q.z; x[6] = q.w;
    out[16 * F + 7] = e->td;
    int k = 0;
    for (auto& it : e->feature)
        if (usable(e, it)) out[16 * F + 8 + k++] = 1. / it.estimated_depth;
    *n_lam = k;
}

// double2vector (estimator.cpp:530-619)
void unpack_states(ve_estimator* e, const double* in) {
    const int F = e->W + 1;
    Vec3 origin_R0 = hm::R2ypr(e->Rs[0]);
    Vec3 origin_P0 = e->Ps[0];
    if (e->failure_occur) {
        origin_R0 = hm::R2ypr(e->last_R0);
        origin_P0 = e->last_P0;
        e->failure_occur = false;
    }
    auto Qp = [&](const double* p) { return Quat(p[6], p[3], p[4], p[5]); };
    const Mat3 R00 = Qp(in).R();
    const Vec3 origin_R00 = hm::R2ypr(R00);
    const double y_diff = origin_R0.x - origin_R00.x;
    Mat3 rot_diff = hm::ypr2R(Vec3(y_diff, 0, 0));
    if (std::fabs(std::fabs(origin_R0.y) - 90) < 1.0 || std::fabs(std::fabs(origin_R00.y) - 90) < 1.0) rot_diff = e->Rs[0] * R00.T();
    for (int i = 0; i < F; i++) {
        const double* p = in + 7 * i;
        const double* s = in + 7 * F + 9 * i;
        e->Rs[i] = rot_diff * Qp(p).normalized().R();
        e->Ps[i] = rot_diff * Vec3(p[0] - in[0], p[1] - in[1], p[2] - in[2]) + origin_P0;
        e->Vs[i] = rot_diff * Vec3(s[0], s[1], s[2]);
        e->Bas[i] = Vec3(s[3], s[4], s[5]);
        e->Bgs[i] = Vec3(s[6], s[7], s[8]);
    }
    const double* x = in + 16 * F;
    e->tic = Vec3(x[0], x[1], x[2]);
    e->ric = Qp(x).R();
    int k = 0;
    for (auto& it : e->feature) {  // FeatureManager::setDepth
        if (!usable(e, it)) continue;
        it.estimated_depth = 1.0 / in[16 * F + 8 + k++];
        it.solve_flag = it.estimated_depth < 0 ? 2 : 1;
    }
    if (e->cfg.estimate_td) e->td = in[16 * F + 7];
}

int upload_prior_meta(ve_estimator* e, int buf) {
    const int nb = (int)e->prior_blocks.size();
    for (int b = 0; b < nb; b++) {
        e->h_prior_i[b] = e->prior_blocks[b].type;
        e->h_prior_i[64 + b] = e->prior_blocks[b].index;
        e->h_prior_i[128 + b] = e->prior_blocks[b].off;
    }
    VE_CUDA(cudaMemcpyAsync(e->d_prior_i[buf].p, e->h_prior_i, sizeof(int) * 192, cudaMemcpyHostToDevice, e->stream));
    return VE_OK;
}

vb::BaPrior prior_view(const ve_estimator* e) {
    vb::BaPrior pr{};
    if (!e->has_prior) return pr;
    const int buf = e->prior_buf;
    const int nm = e->nmax;
    pr.n = e->prior_n;
    pr.nblocks = (int)e->prior_blocks.size();
    pr.type = e->d_prior_i[buf].p;
    pr.index = pr.type + 64;
    pr.off = pr.type + 128;
    pr.A = e->d_prior[buf].p;
    pr.g0 = pr.A + (size_t)nm * nm;
    pr.c0 = pr.g0 + nm;
    pr.x0 = pr.c0 + 1;
    return pr;
}

// Fills a BaProblem for the current window; uploads the observation table and the states to x[0].
int build_problem(ve_estimator* e, vb::BaProblem& p, bool upload_tables) {
    const int W = e->W, F = W + 1;
    int L = 0, M = 0;
    int* lm_anchor = e->h_ints;
    int* lm_start = lm_anchor + e->Lmax;
    int* ob_frame = lm_start + e->Lmax + 1;
    int* imu_slot = ob_frame + e->Mmax;
    double* lm_pts = e->h_obs;
    double* lm_vel = lm_pts + 2 * e->Lmax;
    double* lm_td = lm_vel + 2 * e->Lmax;
    double* lm_row = lm_td + e->Lmax;
    double* ob_pts = lm_row + e->Lmax;
    double* ob_vel = ob_pts + 2 * e->Mmax;
    double* ob_td = ob_vel + 2 * e->Mmax;
    double* ob_row = ob_td + e->Mmax;
    for (auto& it : e->feature) {
        if (!usable(e, it)) continue;
        if (L >= e->Lmax || M + (int)it.feature_per_frame.size() > e->Mmax) {
            e->err = "feature capacity (max_features) exceeded";  // NUM_OF_F overflow is silent in the reference (estimator.h:111)
            return VE_ERR_CAPACITY;
        }
        const FeaturePerFrame& f0 = it.feature_per_frame[0];
        lm_anchor[L] = it.start_frame;
        lm_start[L] = M;
        lm_pts[2 * L] = f0.point.x; lm_pts[2 * L + 1] = f0.point.y;
        lm_vel[2 * L] = f0.vx; lm_vel[2 * L + 1] = f0.vy;
        lm_td[L] = f0.cur_td;
        lm_row[L] = f0.v;
        for (size_t k = 1; k < it.feature_per_frame.size(); k++) {
            const FeaturePerFrame& fj = it.feature_per_frame[k];
            ob_frame[M] = it.start_frame + (int)k;
            ob_pts[2 * M] = fj.point.x; ob_pts[2 * M + 1] = fj.point.y;
            ob_vel[2 * M] = fj.vx; ob_vel[2 * M + 1] = fj.vy;
            ob_td[M] = fj.cur_td;
            ob_row[M] = fj.v;
            M++;
        }
        L++;
    }
    lm_start[L] = M;
    for (int k = 0; k < W; k++) imu_slot[k] = (e->sum_dt[k + 1] > 10.0 || !e->slot_valid[k + 1]) ? -1 : e->slot_of[k + 1];
    if (upload_tables) {
        const size_t nints = (size_t)e->Lmax + (e->Lmax + 1) + e->Mmax + W;
        const size_t nobs = 6 * (size_t)e->Lmax + 6 * (size_t)e->Mmax;
        VE_CUDA(cudaMemcpyAsync(e->d_ints.p, e->h_ints, sizeof(int) * nints, cudaMemcpyHostToDevice, e->stream));
        VE_CUDA(cudaMemcpyAsync(e->d_obs.p, e->h_obs, sizeof(double) * nobs, cudaMemcpyHostToDevice, e->stream));
        e->h2d_bytes += sizeof(int) * nints + sizeof(double) * nobs;
    }
    vb::BaDims& d = p.dims;
    d.W = W;
    d.L = L;
    d.M = M;
    d.est_ex = e->cfg.estimate_extrinsic ? 1 : 0;
    d.est_td = e->cfg.estimate_td ? 1 : 0;
    d.col_sb = 6 * F;
    d.col_ex = d.est_ex ? 15 * F : -1;
    d.col_td = d.est_td ? 15 * F + 6 * d.est_ex : -1;
    d.D = 15 * F + 6 * d.est_ex + d.est_td;
    d.sqrt_info_vis = e->cfg.focal_length / 1.5;
    d.tr_over_row = e->cfg.tr / e->cfg.row;
    d.half_row = e->cfg.row / 2;
    d.G[0] = 0; d.G[1] = 0; d.G[2] = e->cfg.g_norm;
    // the accumulation buffers were sized for the maximum D; the kernels index with dims.D
    for (int b = 0; b < 2; b++) {
        p.x[b] = states_view(e, b);
        double* a = e->d_acc[b].p;
        p.acc[b].Hpp = a;
        p.acc[b].gp = a + (size_t)d.D * d.D;
        p.acc[b].Hpl = p.acc[b].gp + d.D;
        p.acc[b].Hll = p.acc[b].Hpl + (size_t)e->Lmax * d.D;
        p.acc[b].gl = p.acc[b].Hll + e->Lmax;
        p.acc[b].cost = p.acc[b].gl + e->Lmax;
    }
    const int* di = e->d_ints.p;
    p.lm_anchor = di;
    p.lm_start = di + e->Lmax;
    p.ob_frame = di + 2 * e->Lmax + 1;
    p.imu_slot = p.ob_frame + e->Mmax;
    const double* dobs = e->d_obs.p;
    p.lm_pts = dobs;
    p.lm_vel = dobs + 2 * e->Lmax;
    p.lm_td = dobs + 4 * e->Lmax;
    p.lm_row = dobs + 5 * e->Lmax;
    p.ob_pts = dobs + 6 * e->Lmax;
    p.ob_vel = p.ob_pts + 2 * e->Mmax;
    p.ob_td = p.ob_vel + 2 * e->Mmax;
    p.ob_row = p.ob_td + e->Mmax;
    p.preint = e->d_preint.p;
    p.prior = prior_view(e);
    p.S = e->d_S.p;
    p.Spk = e->d_Spk.p;
    p.Hfull = e->d_Hfull.p;
    p.gred = e->d_gred.p;
    const size_t N = (size_t)e->D + e->Lmax;
    p.scale = e->d_vec.p;
    p.diag = p.scale + N;
    p.grad = p.diag + N;
    p.gn = p.grad + N;
    p.work = e->d_work.p;
    p.st = e->d_st.p;
    e->last_landmarks = L;
    e->last_visual = M;
    return VE_OK;
}

int upload_states(ve_estimator* e, int* n_lam) {
    if (e->states_upload_pending) {  // marginalize() of the previous frame staged its linearisation point here
        VE_CUDA(cudaEventSynchronize(e->ev[6]));
        e->states_upload_pending = false;
    }
    pack_states(e, e->h_states, n_lam);
    VE_CUDA(cudaMemcpyAsync(e->d_states[0].p, e->h_states, sizeof(double) * states_doubles(e), cudaMemcpyHostToDevice, e->stream));
    e->h2d_bytes += sizeof(double) * states_doubles(e);
    return VE_OK;
}

int finish_marg(ve_estimator* e);

// The marginalisation branches of Estimator::optimization (estimator.cpp:826-999)
int marginalize(ve_estimator* e, vb::BaProblem& p) {
    const int W = e->W, F = W + 1;
    const bool old = e->marginalization_flag == 0;
    auto prior_has = [&](int type, int index) {
        for (auto& b : e->prior_blocks)
            if (b.type == type && (type >= 2 || b.index == index)) return true;
        return false;
    };
    if (!old && !(e->has_prior && prior_has(0, W - 1))) return VE_OK;
    // the previous marginalisation's diagnostics are collected before its staging area is reused (the stream has just
    // been drained by the solve's read-back: no wait)
    int rc = finish_marg(e);
    if (rc) return rc;
    // vector2double() after double2vector(): the linearisation point is the re-anchored state
    int n_lam = 0;
    rc = upload_states(e, &n_lam);
    if (rc) return rc;
    VE_CUDA(cudaEventRecord(e->ev[6], e->stream));
    e->states_upload_pending = true;
    VE_CUDA(cudaEventRecord(e->ev[5], e->stream));
    // which parameter blocks take part
    std::vector<bool> t_pose(F, false), t_sb(F, false);
    bool t_ex = false, t_td = false;
    if (e->has_prior)
        for (auto& b : e->prior_blocks) {
            if (b.type == 0) t_pose[b.index] = true;
            else if (b.type == 1) t_sb[b.index] = true;
            else if (b.type == 2) t_ex = true;
            else t_td = true;
        }
    vb::MargPlan mp{};
    std::vector<int> lms, col_lm;
    mp.use_imu = 0;
    if (old) {
        if (e->slot_valid[1] && e->sum_dt[1] < 10.0) {
            mp.use_imu = 1;
            t_pose[0] = t_sb[0] = t_pose[1] = t_sb[1] = true;
        }
        int li = -1;
        for (auto& it : e->feature) {
            if (!usable(e, it)) continue;
            ++li;
            if (it.start_frame != 0) continue;
            lms.push_back(li);
            t_pose[0] = true;
            for (size_t k = 1; k < it.feature_per_frame.size(); k++) t_pose[k] = true;
            t_ex = true;
            if (e->cfg.estimate_td) t_td = true;
        }
    }
    // column layout: marginalised dense blocks, marginalised landmarks, kept blocks (canonical order)
    int pos = 0;
    for (int f = 0; f < vb::BA_MAX_FRAMES; f++) mp.col_pose[f] = mp.col_sb[f] = -1;
    mp.col_ex = mp.col_td = -1;
    const int drop_pose = old ? 0 : W - 1;
    if (t_pose[drop_pose]) { mp.col_pose[drop_pose] = pos; pos += 6; }
    if (old && t_sb[0]) { mp.col_sb[0] = pos; pos += 9; }
    mp.m_dense = pos;
    for (size_t k = 0; k < lms.size(); k++) col_lm.push_back(pos++);
    mp.n_lm = (int)lms.size();
    const int m = pos;
    std::vector<PriorBlock> kept;
    for (int f = 0; f < F; f++)
        if (t_pose[f] && mp.col_pose[f] < 0) { mp.col_pose[f] = pos; kept.push_back({0, f, pos - m, 6}); pos += 6; }
    for (int f = 0; f < F; f++)
        if (t_sb[f] && mp.col_sb[f] < 0) { mp.col_sb[f] = pos; kept.push_back({1, f, pos - m, 9}); pos += 9; }
    if (t_ex) { mp.col_ex = pos; kept.push_back({2, 0, pos - m, 6}); pos += 6; }
    if (t_td) { mp.col_td = pos; kept.push_back({3, 0, pos - m, 1}); pos += 1; }
    mp.P = pos;
    mp.n = pos - m;
    if (mp.n > e->nmax || mp.P > e->nmax + 15 + e->Lmax || mp.m_dense == 0) {
        e->err = "marginalisation system larger than the configured capacity";
        return VE_ERR_CAPACITY;
    }
    for (size_t k = 0; k < lms.size(); k++) {
        e->h_marg_i[k] = lms[k];
        e->h_marg_i[e->Lmax + k] = col_lm[k];
    }
    VE_CUDA(cudaMemcpyAsync(e->d_marg_i.p, e->h_marg_i, sizeof(int) * 2 * e->Lmax, cudaMemcpyHostToDevice, e->stream));
    mp.lms = e->d_marg_i.p;
    mp.col_lm = e->d_marg_i.p + e->Lmax;
    const size_t Pm = (size_t)e->nmax + 15 + e->Lmax;
    mp.Am = e->d_marg.p;
    mp.bm = mp.Am + Pm * Pm;
    mp.Araw = mp.bm + Pm;
    mp.graw = mp.Araw + (size_t)e->nmax * e->nmax;
    mp.Wglobal = mp.graw + e->nmax;  // (nmax + 15)^2 doubles
    mp.w_in_global = 0;
    const int nb = e->prior_buf ^ 1;
    mp.Aout = e->d_prior[nb].p;
    mp.gout = mp.Aout + (size_t)e->nmax * e->nmax;
    mp.cout = mp.gout + e->nmax;
    // NOTE: Aout is written with leading dimension n (dense n x n at the front of the buffer)
    vb::launch_marginalize(p, mp, e->stream, &e->last_launches, &e->prof);
    VE_CUDA(cudaGetLastError());
    // new prior: linearisation point = current parameter values of the kept blocks, identities shifted
    // (addr_shift, estimator.cpp:913-925 / :969-990)
    double* x0 = e->h_prior;
    std::vector<PriorBlock> shifted;
    for (size_t b = 0; b < kept.size(); b++) {
        PriorBlock nbk = kept[b];
        double* dst = x0 + 9 * b;
        std::memset(dst, 0, 9 * sizeof(double));
        const double* src = nullptr;
        int gsz = 0;
        if (nbk.type == 0) { src = e->h_states + 7 * nbk.index; gsz = 7; }
        else if (nbk.type == 1) { src = e->h_states + 7 * F + 9 * nbk.index; gsz = 9; }
        else if (nbk.type == 2) { src = e->h_states + 16 * F; gsz = 7; }
        else { src = e->h_states + 16 * F + 7; gsz = 1; }
        std::memcpy(dst, src, gsz * sizeof(double));
        if (nbk.type <= 1) {
            if (old) nbk.index -= 1;
            else if (nbk.index == W) nbk.index = W - 1;
        }
        shifted.push_back(nbk);
    }
    double* d_x0 = e->d_prior[nb].p + (size_t)e->nmax * e->nmax + e->nmax + 1;
    VE_CUDA(cudaMemcpyAsync(d_x0, x0, sizeof(double) * 9 * kept.size(), cudaMemcpyHostToDevice, e->stream));
    // un-thresholded Schur complement for tests
    VE_CUDA(cudaMemcpyAsync(e->h_marg_out, mp.Araw, sizeof(double) * ((size_t)e->nmax * e->nmax + e->nmax), cudaMemcpyDeviceToHost, e->stream));
    e->prior_blocks = shifted;
    e->prior_n = mp.n;
    e->prior_buf = nb;
    e->has_prior = true;
    rc = upload_prior_meta(e, nb);
    if (rc) return rc;
    // No host wait: the new prior is consumed by the next solve on the same (in-order) stream, so the marginalisation
    // overlaps whatever the caller does next (the reference's tracker node runs in another process anyway).
    e->marg_pending = true;
    e->marg_n = mp.n;
    return VE_OK;
}

// Waits for a pending marginalisation and collects its diagnostics (Schur complement before the eps floor).
int finish_marg(ve_estimator* e) {
    if (!e->marg_pending) return VE_OK;
    VE_CUDA(cudaEventSynchronize(e->ev[3]));
    e->marg_pending = false;
    const int n = e->marg_n;
    e->prior_raw_A.assign(e->h_marg_out, e->h_marg_out + (size_t)n * n);
    e->prior_raw_b.assign(e->h_marg_out + (size_t)e->nmax * e->nmax, e->h_marg_out + (size_t)e->nmax * e->nmax + n);
    if (n + 7 <= e->nmax)
        for (int k = 0; k < 7; k++) e->marg_sweeps[k] = e->h_marg_out[(size_t)e->nmax * e->nmax + n + k];
    cudaEventElapsedTime(&e->last_ms[2], e->ev[5], e->ev[3]);
    return VE_OK;
}

int optimization(ve_estimator* e) {
    int rc;
    // No wait for the previous frame's marginalisation here: everything below is enqueued behind it on the same stream
    // while it is still running, so the host side of a frame (bookkeeping, problem tables, uploads, 26 launches) is
    // hidden instead of sitting between two kernels.
    VE_CUDA(cudaEventRecord(e->ev[0], e->stream));
    for (int f = 0; f <= e->W; f++)
        if ((rc = flush_frame(e, f))) return rc;
    if ((rc = refresh_sqrt_info(e))) return rc;
    VE_CUDA(cudaEventRecord(e->ev[1], e->stream));
    vb::BaProblem p{};
    if ((rc = build_problem(e, p, true))) return rc;
    int n_lam = 0;
    if ((rc = upload_states(e, &n_lam))) return rc;
    vb::SolverState& st = *e->h_st;
    std::memset(&st, 0, sizeof(st));
    st.max_iterations = e->cfg.num_iterations;
    st.first = 1;
    st.radius = 1e4;
    st.mu = 1e-8;
    VE_CUDA(cudaMemcpyAsync(e->d_st.p, &st, sizeof(st), cudaMemcpyHostToDevice, e->stream));
    vb::launch_ba_solve(p, e->cfg.num_iterations, e->stream, &e->last_launches, &e->prof);
    VE_CUDA(cudaGetLastError());
    VE_CUDA(cudaMemcpyAsync(&st, e->d_st.p, sizeof(st), cudaMemcpyDeviceToHost, e->stream));
    const size_t ns = states_doubles(e);
    double* h2 = e->h_states + ns;  // second half of the staging area holds both device buffers
    VE_CUDA(cudaMemcpyAsync(h2, e->d_states[0].p, sizeof(double) * ns, cudaMemcpyDeviceToHost, e->stream));
    VE_CUDA(cudaMemcpyAsync(h2 + ns, e->d_states[1].p, sizeof(double) * ns, cudaMemcpyDeviceToHost, e->stream));
    e->d2h_bytes += 2 * sizeof(double) * ns + sizeof(st);
    VE_CUDA(cudaEventRecord(e->ev[2], e->stream));
    VE_CUDA(cudaStreamSynchronize(e->stream));
    e->flushes_in_flight = 0;
    e->last_state = st;
    e->n_solves++;
    unpack_states(e, h2 + (size_t)st.cur * ns);
    cudaEventElapsedTime(&e->last_ms[0], e->ev[0], e->ev[1]);
    cudaEventElapsedTime(&e->last_ms[1], e->ev[1], e->ev[2]);
    cudaEventElapsedTime(&e->last_ms[3], e->ev[0], e->ev[2]);
    rc = marginalize(e, p);
    if (rc) return rc;
    VE_CUDA(cudaEventRecord(e->ev[3], e->stream));
    if (e->prof.on && (rc = finish_marg(e))) return rc;  // profiling runs are serialised anyway
    return VE_OK;
}

int solve_odometry(ve_estimator* e) {
    if (e->frame_count < e->W) return VE_OK;
    if (e->solver_flag == 1) {
        triangulate(e);
        return optimization(e);
    }
    return VE_OK;
}

void slide_window_old(ve_estimator* e) {
    if (e->solver_flag == 1) {
        const Mat3 R0 = e->back_R0 * e->ric, R1 = e->Rs[0] * e->ric;
        const Vec3 P0 = e->back_P0 + e->back_R0 * e->tic, P1 = e->Ps[0] + e->Rs[0] * e->tic;
        remove_back_shift_depth(e, R0, P0, R1, P1);
    } else
        remove_back(e);
}

int slide_window(ve_estimator* e) {
    const int W = e->W;
    if (e->marginalization_flag == 0) {
        e->back_R0 = e->Rs[0];
        e->back_P0 = e->Ps[0];
        if (e->frame_count == W) {
            const int slot0 = e->slot_of[0];
            for (int i = 0; i < W; i++) {
                std::swap(e->Rs[i], e->Rs[i + 1]);
                e->slot_of[i] = e->slot_of[i + 1];
                { bool t = e->slot_valid[i]; e->slot_valid[i] = e->slot_valid[i + 1]; e->slot_valid[i + 1] = t; }
                std::swap(e->flushed[i], e->flushed[i + 1]);
                std::swap(e->sum_dt[i], e->sum_dt[i + 1]);
                std::swap(e->lin_acc[i], e->lin_acc[i + 1]);
                std::swap(e->lin_gyr[i], e->lin_gyr[i + 1]);
                e->dt_buf[i].swap(e->dt_buf[i + 1]);
                e->acc_buf[i].swap(e->acc_buf[i + 1]);
                e->gyr_buf[i].swap(e->gyr_buf[i + 1]);
                e->Headers[i] = e->Headers[i + 1];
                std::swap(e->Ps[i], e->Ps[i + 1]);
                std::swap(e->Vs[i], e->Vs[i + 1]);
                std::swap(e->Bas[i], e->Bas[i + 1]);
                std::swap(e->Bgs[i], e->Bgs[i + 1]);
            }
            e->slot_of[W] = slot0;
            e->Headers[W] = e->Headers[W - 1];
            e->Ps[W] = e->Ps[W - 1]; e->Vs[W] = e->Vs[W - 1]; e->Rs[W] = e->Rs[W - 1];
            e->Bas[W] = e->Bas[W - 1]; e->Bgs[W] = e->Bgs[W - 1];
            const int rc = init_slot(e, W, e->acc_0, e->gyr_0, e->Bas[W], e->Bgs[W]);
            if (rc) return rc;
            e->dt_buf[W].clear(); e->acc_buf[W].clear(); e->gyr_buf[W].clear();
            slide_window_old(e);
        }
    } else if (e->frame_count == W) {
        const int fc = e->frame_count;
        for (size_t i = 0; i < e->dt_buf[fc].size(); i++) {
            e->dt_buf[fc - 1].push_back(e->dt_buf[fc][i]);
            e->acc_buf[fc - 1].push_back(e->acc_buf[fc][i]);
            e->gyr_buf[fc - 1].push_back(e->gyr_buf[fc][i]);
            e->sum_dt[fc - 1] += e->dt_buf[fc][i];
        }
        int rc = flush_frame(e, fc - 1);  // pre_integrations[frame_count - 1]->push_back(...)
        if (rc) return rc;
        e->Headers[fc - 1] = e->Headers[fc];
        e->Ps[fc - 1] = e->Ps[fc]; e->Vs[fc - 1] = e->Vs[fc]; e->Rs[fc - 1] = e->Rs[fc];
        e->Bas[fc - 1] = e->Bas[fc]; e->Bgs[fc - 1] = e->Bgs[fc];
        rc = init_slot(e, W, e->acc_0, e->gyr_0, e->Bas[W], e->Bgs[W]);
        if (rc) return rc;
        e->dt_buf[W].clear(); e->acc_buf[W].clear(); e->gyr_buf[W].clear();
        remove_front(e, fc);
    }
    return VE_OK;
}

void remove_failures(ve_estimator* e) {
    e->feature.erase(std::remove_if(e->feature.begin(), e->feature.end(), [](const FeaturePerId& f) { return f.solve_flag == 2; }),
                     e->feature.end());
}

}  // namespace

extern "C" {

int ve_create(const ve_config* cfg, ve_estimator** out) {
    if (!cfg || !out) return VE_ERR_INVALID;
    *out = nullptr;
    if (cfg->window_size < 3 || cfg->window_size + 1 > vb::BA_MAX_FRAMES || cfg->window_size + 1 > vb::BA_MAX_OBS_PER_LM ||
        cfg->max_features < 8 || cfg->num_iterations < 1 || cfg->estimate_extrinsic > 1)
        return VE_ERR_INVALID;
    // work arrays of the single-CTA solvers: prior of 6 W + 9 + 6 + 1 <= 160 parameters, reduced system of
    // 15 (W + 1) + 7 <= 352 columns, i.e. WINDOW_SIZE <= 22 (the reference ships 10; above 13 the reduced systems
    // move from shared to global memory)
    if (6 * cfg->window_size + 16 > 160 || 15 * (cfg->window_size + 1) + 7 > 352) return VE_ERR_INVALID;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || cfg->device < 0 || cfg->device >= ndev) return VE_ERR_NO_DEVICE;
    ve_estimator* e = new ve_estimator();
    e->cfg = *cfg;
    e->W = cfg->window_size;
    const int W = e->W, F = W + 1;
    e->Ps.resize(F); e->Vs.resize(F); e->Bas.resize(F); e->Bgs.resize(F); e->Rs.resize(F);
    e->Headers.assign(F, 0.0);
    e->dt_buf.resize(F); e->acc_buf.resize(F); e->gyr_buf.resize(F);
    e->slot_of.resize(F); e->slot_valid.assign(F, false); e->flushed.assign(F, 0); e->sum_dt.assign(F, 0.0);
    e->lin_acc.resize(F); e->lin_gyr.resize(F); e->sqrt_dirty.assign(F, true);
    e->Lmax = cfg->max_features;
    e->Mmax = cfg->max_features * W;
    e->D = 15 * F + 6 + 1;
    e->nmax = 6 * F + 9 * 2 + 6 + 1;
    clear_state(e);
    set_parameter(e);
    auto fail = [&](int code) {
        ve_destroy(e);
        return code;
    };
#define VE_TRY(call)                                          \
    do {                                                      \
        if ((call) != cudaSuccess) return fail(VE_ERR_CUDA);  \
    } while (0)
    VE_TRY(cudaSetDevice(cfg->device));
    VE_TRY(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
    for (auto& ev : e->ev) VE_TRY(cudaEventCreate(&ev));
    VE_TRY(e->d_preint.alloc(F));
    VE_TRY(e->d_samples.alloc(8 * 7 * 512));
    VE_TRY(e->d_which.alloc(F));
    const size_t Pm = (size_t)e->nmax + 15 + e->Lmax;
    for (int b = 0; b < 2; b++) {
        VE_TRY(e->d_states[b].alloc(states_doubles(e)));
        VE_TRY(e->d_acc[b].alloc(acc_doubles(e)));
        VE_TRY(e->d_prior[b].alloc((size_t)e->nmax * e->nmax + e->nmax + 1 + 9 * 64));
        VE_TRY(e->d_prior_i[b].alloc(192));
    }
    VE_TRY(e->d_ints.alloc((size_t)2 * e->Lmax + 1 + e->Mmax + W));
    VE_TRY(e->d_obs.alloc(6 * (size_t)e->Lmax + 6 * (size_t)e->Mmax));
    VE_TRY(e->d_S.alloc((size_t)e->D * e->D));
    VE_TRY(e->d_Spk.alloc((size_t)e->D * (e->D + 1) / 2));
    VE_TRY(e->d_Hfull.alloc((size_t)e->D * e->D));
    VE_TRY(e->d_gred.alloc(e->D));
    VE_TRY(e->d_vec.alloc(4 * ((size_t)e->D + e->Lmax)));
    VE_TRY(e->d_work.alloc(vb::ba_work_doubles(e->D, e->Lmax)));
    VE_TRY(e->d_st.alloc(1));
    VE_TRY(e->d_marg.alloc(Pm * Pm + Pm + (size_t)e->nmax * e->nmax + e->nmax + ((size_t)e->nmax + 15) * (e->nmax + 15)));
    VE_TRY(e->d_marg_i.alloc(2 * (size_t)e->Lmax));
    VE_TRY(cudaHostAlloc(&e->h_states, sizeof(double) * 3 * states_doubles(e), cudaHostAllocDefault));
    VE_TRY(cudaHostAlloc(&e->h_obs, sizeof(double) * (6 * (size_t)e->Lmax + 6 * (size_t)e->Mmax), cudaHostAllocDefault));
    VE_TRY(cudaHostAlloc(&e->h_ints, sizeof(int) * ((size_t)2 * e->Lmax + 1 + e->Mmax + W), cudaHostAllocDefault));
    VE_TRY(cudaHostAlloc(&e->h_samples, sizeof(double) * 8 * 7 * 512, cudaHostAllocDefault));
    VE_TRY(cudaHostAlloc(&e->h_preint, sizeof(vb::PreInt), cudaHostAllocDefault));
    VE_TRY(cudaHostAlloc(&e->h_st, sizeof(vb::SolverState), cudaHostAllocDefault));
    VE_TRY(cudaHostAlloc(&e->h_prior, sizeof(double) * 9 * 64, cudaHostAllocDefault));
    VE_TRY(cudaHostAlloc(&e->h_prior_i, sizeof(int) * 192, cudaHostAllocDefault));
    VE_TRY(cudaHostAlloc(&e->h_marg_i, sizeof(int) * std::max<size_t>(2 * (size_t)e->Lmax, 64), cudaHostAllocDefault));
    VE_TRY(cudaHostAlloc(&e->h_which, sizeof(int) * 64, cudaHostAllocDefault));
    VE_TRY(cudaHostAlloc(&e->h_marg_out, sizeof(double) * ((size_t)e->nmax * e->nmax + e->nmax), cudaHostAllocDefault));
#undef VE_TRY
    *out = e;
    return VE_OK;
}

void ve_destroy(ve_estimator* e) {
    if (!e) return;
    cudaSetDevice(e->cfg.device);
    if (e->stream) cudaStreamSynchronize(e->stream);
    e->d_preint.release(); e->d_samples.release(); e->d_which.release();
    for (int b = 0; b < 2; b++) { e->d_states[b].release(); e->d_acc[b].release(); e->d_prior[b].release(); e->d_prior_i[b].release(); }
    e->d_ints.release(); e->d_obs.release(); e->d_S.release(); e->d_Spk.release(); e->d_Hfull.release(); e->d_gred.release(); e->d_vec.release();
    e->d_work.release(); e->d_st.release(); e->d_marg.release(); e->d_marg_i.release();
    cudaFreeHost(e->h_states); cudaFreeHost(e->h_obs); cudaFreeHost(e->h_ints); cudaFreeHost(e->h_samples); cudaFreeHost(e->h_preint);
    cudaFreeHost(e->h_st); cudaFreeHost(e->h_prior); cudaFreeHost(e->h_prior_i); cudaFreeHost(e->h_marg_i); cudaFreeHost(e->h_which); cudaFreeHost(e->h_marg_out);
    for (auto& ev : e->ev)
        if (ev) cudaEventDestroy(ev);
    if (e->stream) cudaStreamDestroy(e->stream);
    delete e;
}

const char* ve_last_error(const ve_estimator* e) { return e ? e->err.c_str() : "null handle"; }

int ve_clear_state(ve_estimator* e) {
    if (!e) return VE_ERR_INVALID;
    clear_state(e);
    set_parameter(e);
    return VE_OK;
}

int ve_set_seed(ve_estimator* e, int n, const double* rows, const double* ba, const double* bg) {
    if (!e || n < 0 || (n && !rows) || !ba || !bg) return VE_ERR_INVALID;
    e->seeds.clear();
    for (int i = 0; i < n; i++) {
        const double* r = rows + 11 * i;
        SeedRow s;
        s.t = r[0];
        s.P = Vec3(r[1], r[2], r[3]);
        s.R = Quat(r[4], r[5], r[6], r[7]).normalized().R();
        s.V = Vec3(r[8], r[9], r[10]);
        e->seeds.push_back(s);
    }
    e->seed_ba = Vec3(ba[0], ba[1], ba[2]);
    e->seed_bg = Vec3(bg[0], bg[1], bg[2]);
    return VE_OK;
}

int ve_process_imu(ve_estimator* e, double dt, const double* acc, const double* gyr) {
    if (!e || !acc || !gyr) return VE_ERR_INVALID;
    VE_CUDA(cudaSetDevice(e->cfg.device));
    return process_imu(e, dt, Vec3(acc[0], acc[1], acc[2]), Vec3(gyr[0], gyr[1], gyr[2]));
}

int ve_process_image(ve_estimator* e, int n, const int* ids, const double* xyz_uv_vel, double stamp) {
    if (!e || n < 0 || (n && (!ids || !xyz_uv_vel))) return VE_ERR_INVALID;
    VE_CUDA(cudaSetDevice(e->cfg.device));
    e->last_launches = 0;
    e->h2d_bytes = e->d2h_bytes = 0;
    e->last_ms[0] = e->last_ms[1] = e->last_ms[2] = e->last_ms[3] = 0;
    std::vector<int> order(n);
    for (int i = 0; i < n; i++) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return ids[a] < ids[b]; });
    e->marginalization_flag = add_feature_check_parallax(e, n, order.data(), ids, xyz_uv_vel, e->td) ? 0 : 1;
    e->Headers[e->frame_count] = stamp;
    int rc = VE_OK;
    const int W = e->W;
    if (e->solver_flag == 0) {
        if (e->frame_count == W) {
            bool ok = false;
            if ((rc = initial_from_seed(e, &ok))) return rc;
            if (ok) {
                e->solver_flag = 1;
                if ((rc = solve_odometry(e))) return rc;
                if ((rc = slide_window(e))) return rc;
                remove_failures(e);
                e->last_R = e->Rs[W]; e->last_P = e->Ps[W]; e->last_R0 = e->Rs[0]; e->last_P0 = e->Ps[0];
            } else if ((rc = slide_window(e)))
                return rc;
        } else
            e->frame_count++;
    } else {
        if ((rc = solve_odometry(e))) return rc;
        if (failure_detection(e)) {
            e->failure_occur = true;
            clear_state(e);
            set_parameter(e);
            e->n_reboots++;
            return VE_OK;
        }
        if ((rc = slide_window(e))) return rc;
        remove_failures(e);
        e->last_R = e->Rs[W]; e->last_P = e->Ps[W]; e->last_R0 = e->Rs[0]; e->last_P0 = e->Ps[0];
    }
    return VE_OK;
}

int ve_get_states(const ve_estimator* e, double* out, double* td) {
    if (!e || !out) return VE_ERR_INVALID;
    for (int i = 0; i <= e->W; i++) {
        double* o = out + 16 * i;
        const Quat q = Quat::FromR(e->Rs[i]);
        o[0] = e->Ps[i].x; o[1] = e->Ps[i].y; o[2] = e->Ps[i].z; o[3] = q.w; o[4] = q.x; o[5] = q.y; o[6] = q.z;
        o[7] = e->Vs[i].x; o[8] = e->Vs[i].y; o[9] = e->Vs[i].z;
        o[10] = e->Bas[i].x; o[11] = e->Bas[i].y; o[12] = e->Bas[i].z;
        o[13] = e->Bgs[i].x; o[14] = e->Bgs[i].y; o[15] = e->Bgs[i].z;
    }
    if (td) *td = e->td;
    return VE_OK;
}

int ve_info(const ve_estimator* e, int* o, double* costs2) {
    if (!e || !o) return VE_ERR_INVALID;
    o[0] = e->solver_flag; o[1] = e->frame_count; o[2] = e->marginalization_flag; o[3] = e->n_solves; o[4] = e->n_reboots;
    o[5] = e->last_landmarks; o[6] = e->last_visual; o[7] = e->last_state.iteration; o[8] = e->last_state.successful;
    // termination coding: 0 iteration cap, 1 parameter tol, 2 function tol, 4 failure
    o[9] = e->last_state.done == 2 ? 1 : e->last_state.done == 3 ? 2 : e->last_state.done == 4 ? 4 : 0;
    if (costs2) {
        costs2[0] = e->last_state.initial_cost;
        costs2[1] = e->last_state.x_cost;
    }
    return VE_OK;
}

int ve_get_prior(const ve_estimator* ce, int cap, double* A, double* b, int* nblocks, int* blocks4) {
    if (!ce) return VE_ERR_INVALID;
    ve_estimator* e = const_cast<ve_estimator*>(ce);
    if (!e->has_prior) return 0;
    if (finish_marg(e)) return VE_ERR_CUDA;
    const int n = e->prior_n;
    if (n > cap) return -n;
    std::memcpy(A, e->prior_raw_A.data(), sizeof(double) * (size_t)n * n);
    std::memcpy(b, e->prior_raw_b.data(), sizeof(double) * n);
    if (nblocks) *nblocks = (int)e->prior_blocks.size();
    if (blocks4)
        for (size_t k = 0; k < e->prior_blocks.size(); k++) {
            blocks4[4 * k] = e->prior_blocks[k].type;
            blocks4[4 * k + 1] = e->prior_blocks[k].index;
            blocks4[4 * k + 2] = e->prior_blocks[k].off;
            blocks4[4 * k + 3] = e->prior_blocks[k].size;
        }
    return n;
}

int ve_process_imu_batch(ve_estimator* e, int n, const double* dt, const double* acc, const double* gyr) {
    if (!e || n < 0 || (n && (!dt || !acc || !gyr))) return VE_ERR_INVALID;
    VE_CUDA(cudaSetDevice(e->cfg.device));
    for (int i = 0; i < n; i++) {
        const int rc = process_imu(e, dt[i], Vec3(acc[3 * i], acc[3 * i + 1], acc[3 * i + 2]), Vec3(gyr[3 * i], gyr[3 * i + 1], gyr[3 * i + 2]));
        if (rc) return rc;
    }
    return VE_OK;
}

int ve_set_profile(ve_estimator* e, int on) {
    if (!e) return VE_ERR_INVALID;
    e->prof.enable(on != 0);
    return VE_OK;
}

int ve_kernel_times(const ve_estimator* e, double* ms8, int* count8) {
    if (!e) return VE_ERR_INVALID;
    for (int k = 0; k < 8; k++) {
        if (ms8) ms8[k] = e->prof.ms[k];
        if (count8) count8[k] = e->prof.count[k];
    }
    return VE_OK;
}

int ve_last_traffic(const ve_estimator* e, double* h2d_bytes, double* d2h_bytes) {
    if (!e) return VE_ERR_INVALID;
    if (h2d_bytes) *h2d_bytes = (double)e->h2d_bytes;
    if (d2h_bytes) *d2h_bytes = (double)e->d2h_bytes;
    return VE_OK;
}

int ve_solver_debug(const ve_estimator* ce, double* out13) {  // 18 doubles
    if (!ce || !out13) return VE_ERR_INVALID;
    ve_estimator* e = const_cast<ve_estimator*>(ce);
    if (finish_marg(e)) return VE_ERR_CUDA;
    out13[0] = e->last_state.retries;
    out13[1] = e->last_state.mu;
    out13[2] = e->last_state.radius;
    for (int k = 0; k < 8; k++) out13[3 + k] = (double)e->last_state.clk[k];
    out13[11] = e->marg_sweeps[0];
    out13[12] = e->marg_sweeps[1];
    for (int k = 0; k < 5; k++) out13[13 + k] = e->marg_sweeps[2 + k];
    return VE_OK;
}

int ve_last_timing(const ve_estimator* ce, float* ms4, int* launches) {
    if (!ce) return VE_ERR_INVALID;
    ve_estimator* e = const_cast<ve_estimator*>(ce);
    if (ms4 && finish_marg(e)) return VE_ERR_CUDA;
    if (ms4) std::memcpy(ms4, e->last_ms, sizeof(e->last_ms));
    if (launches) *launches = e->last_launches;
    return VE_OK;
}

}  // extern "C"

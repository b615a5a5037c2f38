// Whole-frame DiMP tracking: uint8 frame in, bounding box out (DiMP.track, pytracking/tracker/dimp/dimp.py:94-175).
//
//   host  plan_crop   get_centered_sample_pos (dimp.py:184-188) + the geometry half of sample_patch
//                     (pytracking/features/preprocessing.py:55-148) + get_sample_location (dimp.py:177-182)
//   device            sample_patch_kernel (decimate + replicate pad + bilinear resize, bit-exact with torch-CPU F.interpolate),
//                     backbone + head + classify (net.cu, corr_api.cu), localize_kernel (dimp.py:196-303)
//   host  commit      translation, update_state (dimp.py:486-497), get_iounet_box (:500-507), update_sample_weights (:445-484),
//                     the iteration schedule of update_classifier (:605-625), the output box (:163-171)
//   device            memory insert + steepest-descent update (dimp_state.cu)
//
// The tracker's scalars live on the host as float32 and every expression below performs the float32 operation sequence torch
// executes for the reference's tensor expression (scalars multiplying a float32 tensor are first rounded to float32; `x / tensor`
// is `tensor.reciprocal() * x`, torch/_tensor.py __rdiv__; integer tensors divided by Python ints become float32; torch.round and
// Python round are round-half-even).  The file is compiled with -fmad=false / -ffp-contract=off: no fused multiply-adds except
// the explicit ones of the resampling kernel.
#include "dimp_state.cuh"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>

using namespace b200trk;

#include "dimp_tracker_kernels.cuh"      // sample_patch_kernel, localize_kernel, LocArgs (anonymous namespace)

namespace {

inline float f32(double x) { return (float)x; }

}  // namespace

// ---------------------------------------------------------------------------------------------------------------------------
// host state
// ---------------------------------------------------------------------------------------------------------------------------
struct b200trk_dimp_tracker {
    b200trk_dimp_state* st = nullptr;
    b200trk_dimp_params_t p;
    // DiMP scalars (float32 tensors in the reference), [0] = row / height, [1] = column / width
    float pos[2] = {0, 0}, target_sz[2] = {0, 0}, base_target_sz[2] = {0, 0}, image_sz[2] = {0, 0};
    float img_sample_sz[2] = {0, 0}, feature_sz[2] = {18, 18}, kernel_size[2] = {4, 4}, score_sz[2] = {19, 19};
    float target_scale = 1.f, min_scale_factor = 0.f, max_scale_factor = 0.f;
    int frame_num = 0, initialized = 0;
    // memory bookkeeping (dimp.py:410-484)
    std::vector<float> sw;
    long num_stored = 0; int num_init = 0, prev_replace_ind = -1;
    // IoUNet refinement (dimp.py:650-723)
    b200trk_iou_predictor_t* iou_pred = nullptr;
    float *mod3 = nullptr, *mod4 = nullptr, *iou3 = nullptr, *iou4 = nullptr, *boxes_dev = nullptr;   // device
    float* boxes_host = nullptr;                     // pinned: [16][4] boxes + [16] IoUs
    int iou_dims[6] = {0, 0, 0, 0, 0, 0};
    float pos_iounet[2] = {0, 0}; int has_pos_iounet = 0;
    std::vector<float> noise;                        // uniform numbers for the next frame's random proposals
    uint64_t rng = 0x9E3779B97F4A7C15ull;
    // device side
    cudaStream_t copy_stream = nullptr; cudaEvent_t ev_img = nullptr;   // frame upload overlaps the previous frame's filter update
    uint8_t* img_dev = nullptr; size_t img_cap = 0;
    b200trk_loc_result_t* loc_dev = nullptr;
    b200trk_loc_result_t* loc_host = nullptr;      // pinned
};

// sample_patch geometry (preprocessing.py:55-148, mode 'replicate')
static void plan_patch(const float pos[2], const float sample_sz[2], const float output_sz[2], int H, int W, b200trk_crop_geom_t* g) {
    long posl[2] = {(long)std::trunc(pos[0]), (long)std::trunc(pos[1])};                       // pos.long()
    const float rf = std::fmin(sample_sz[0] / output_sz[0], sample_sz[1] / output_sz[1]);      // torch.min(sample_sz / output_sz).item()
    int df = (int)std::trunc((double)rf - 0.1);                                                // int(resize_factor - 0.1)
    if (df < 1) df = 1;
    float sz[2] = {sample_sz[0] / (float)df, sample_sz[1] / (float)df};
    float poslf[2];
    int os[2] = {0, 0};
    for (int i = 0; i < 2; ++i) {
        if (df > 1) {
            long m = posl[i] % df; if (m < 0) m += df;                                         // torch remainder: sign of the divisor
            os[i] = (int)m;
            poslf[i] = (float)(posl[i] - m) / (float)df;                                       // (posl - os) / df: float32 tensor
        } else {
            poslf[i] = (float)posl[i];
        }
    }
    float tl[2], br[2]; int tli[2], bri[2];
    for (int i = 0; i < 2; ++i) {
        const long szl = (long)std::fmax(std::nearbyint(sz[i]), 2.0f);                         // torch.max(sz.round(), 2).long()
        tl[i] = poslf[i] - ((float)(szl - 1) / 2.0f);                                          // posl - (szl - 1)/2
        br[i] = (poslf[i] + ((float)szl / 2.0f)) + 1.0f;                                       // posl + szl/2 + 1
        tli[i] = (int)std::trunc(tl[i]); bri[i] = (int)std::trunc(br[i]);                      // .int().item()
    }
    (void)H; (void)W;
    g->df = df; g->os_r = os[0]; g->os_c = os[1]; g->tl_r = tli[0]; g->tl_c = tli[1];
    g->in_h = bri[0] - tli[0]; g->in_w = bri[1] - tli[1];
    g->out_h = (int)output_sz[0]; g->out_w = (int)output_sz[1]; g->win_r = 0; g->win_c = 0;
    g->coord[0] = (float)df * tl[0]; g->coord[1] = (float)df * tl[1]; g->coord[2] = (float)df * br[0]; g->coord[3] = (float)df * br[1];
}

// DiMP.get_sample_location (dimp.py:177-182)
static void sample_location(const b200trk_dimp_tracker* t, b200trk_crop_geom_t* g) {
    for (int i = 0; i < 2; ++i) g->sample_pos[i] = 0.5f * ((g->coord[i] + g->coord[2 + i]) - 1.0f);
    const float a = (g->coord[2] - g->coord[0]) / t->img_sample_sz[0], b = (g->coord[3] - g->coord[1]) / t->img_sample_sz[1];
    g->sample_scale = std::sqrt(a * b);
}

// DiMP.get_iounet_box (dimp.py:500-507) -> (x, y, w, h) in crop pixels
static void iounet_box(const b200trk_dimp_tracker* t, const float pos[2], const float sz[2], const float sample_pos[2], float sample_scale,
                       float box[4]) {
    float ul[2], bs[2];
    for (int i = 0; i < 2; ++i) {
        const float bc = (pos[i] - sample_pos[i]) / sample_scale + (t->img_sample_sz[i] - 1.0f) / 2.0f;
        bs[i] = sz[i] / sample_scale;
        ul[i] = bc - (bs[i] - 1.0f) / 2.0f;
    }
    box[0] = ul[1]; box[1] = ul[0]; box[2] = bs[1]; box[3] = bs[0];
}

extern "C" int b200trk_dimp_tracker_create(b200trk_dimp_tracker_t** out, b200trk_dimp_state_t* state, const b200trk_dimp_params_t* params) {
    B200_REQUIRE(out && params, "dimp_tracker_create: null pointer");
    B200_REQUIRE(params->image_sample_size > 0 && params->sample_memory_size >= 1, "dimp_tracker_create: bad parameters");
    B200_REQUIRE(!state || state->memory_size == params->sample_memory_size, "dimp_tracker_create: sample_memory_size=%d but the state holds %d",
                 params->sample_memory_size, state ? state->memory_size : 0);
    B200_REQUIRE(!state || (state->net->crop_h == params->image_sample_size && state->net->crop_w == params->image_sample_size),
                 "dimp_tracker_create: the network was built for %dx%d crops, image_sample_size=%d", state ? state->net->crop_h : 0,
                 state ? state->net->crop_w : 0, params->image_sample_size);
    b200trk_dimp_tracker* t = new b200trk_dimp_tracker();
    t->st = state; t->p = *params;
    t->img_sample_sz[0] = t->img_sample_sz[1] = (float)params->image_sample_size;
    if (state) {
        t->feature_sz[0] = (float)state->Hc; t->feature_sz[1] = (float)state->Wc;
        t->kernel_size[0] = t->kernel_size[1] = (float)state->ksz;
        t->score_sz[0] = (float)state->Ho; t->score_sz[1] = (float)state->Wo;
        if (cudaMalloc((void**)&t->loc_dev, sizeof(b200trk_loc_result_t)) != cudaSuccess ||
            cudaMallocHost((void**)&t->loc_host, sizeof(b200trk_loc_result_t)) != cudaSuccess) {
            set_error("dimp_tracker_create: allocation failed");
            b200trk_dimp_tracker_destroy(t);
            return 1;
        }
    } else {
        t->feature_sz[0] = t->feature_sz[1] = (float)(params->image_sample_size / 16);
        t->score_sz[0] = t->score_sz[1] = t->feature_sz[0] + 1.f;
    }
    t->sw.assign(params->sample_memory_size, 0.f);
    *out = t;
    return 0;
}

extern "C" int b200trk_dimp_tracker_destroy(b200trk_dimp_tracker_t* t) {
    if (!t) return 0;
    if (t->img_dev) cudaFree(t->img_dev);
    if (t->copy_stream) cudaStreamDestroy(t->copy_stream);
    if (t->ev_img) cudaEventDestroy(t->ev_img);
    if (t->loc_dev) cudaFree(t->loc_dev);
    if (t->loc_host) cudaFreeHost(t->loc_host);
    for (float* q : {t->mod3, t->mod4, t->iou3, t->iou4, t->boxes_dev}) if (q) cudaFree(q);
    if (t->boxes_host) cudaFreeHost(t->boxes_host);
    delete t;
    return 0;
}

extern "C" int b200trk_dimp_tracker_init_state(b200trk_dimp_tracker_t* t, int H, int W, const double b[4], b200trk_crop_geom_t* init_crop,
                                               float init_target_box[4]) {
    B200_REQUIRE(t && b, "dimp_tracker_init_state: null pointer");
    B200_REQUIRE(H > 0 && W > 0 && b[2] > 0 && b[3] > 0, "dimp_tracker_init_state: bad image / box");
    // dimp.py:44-76
    t->frame_num = 1;
    t->pos[0] = f32(b[1] + (b[3] - 1) / 2); t->pos[1] = f32(b[0] + (b[2] - 1) / 2);
    t->target_sz[0] = f32(b[3]); t->target_sz[1] = f32(b[2]);
    t->image_sz[0] = (float)H; t->image_sz[1] = (float)W;
    const float sas = f32(t->p.search_area_scale);
    const double search_area = (double)((t->target_sz[0] * sas) * (t->target_sz[1] * sas));       // torch.prod(target_sz * scale).item()
    const float root = std::sqrt(t->img_sample_sz[0] * t->img_sample_sz[1]);                      // img_sample_sz.prod().sqrt()
    t->target_scale = (1.0f / root) * f32(std::sqrt(search_area));                                // float / tensor = reciprocal * float
    for (int i = 0; i < 2; ++i) t->base_target_sz[i] = t->target_sz[i] / t->target_scale;
    t->min_scale_factor = std::fmax((1.0f / t->base_target_sz[0]) * 10.0f, (1.0f / t->base_target_sz[1]) * 10.0f);
    t->max_scale_factor = std::fmin(t->image_sz[0] / t->base_target_sz[0], t->image_sz[1] / t->base_target_sz[1]);
    // generate_init_samples (dimp.py:331-398), un-augmented: Identity transform on the expanded patch
    const float init_pos[2] = {std::nearbyint(t->pos[0]), std::nearbyint(t->pos[1])};               // self.pos.round()
    const float init_scale = t->target_scale;
    float aug_sz[2] = {t->img_sample_sz[0], t->img_sample_sz[1]};
    const double ef = t->p.augmentation_expansion_factor;
    if (ef != 0.0 && ef != 1.0) {
        for (int i = 0; i < 2; ++i) {
            long a = (long)std::trunc(t->img_sample_sz[i] * f32(ef));
            a += (a - (long)t->img_sample_sz[i]) % 2;
            aug_sz[i] = (float)a;
        }
    }
    const float sample_sz[2] = {init_scale * aug_sz[0], init_scale * aug_sz[1]};
    b200trk_crop_geom_t g;
    plan_patch(init_pos, sample_sz, aug_sz, H, W, &g);
    // Identity.crop_to_output (augmentation.py:20-37): pad by (output - image)/2, floor on the top / left side
    g.win_r = -(int)std::floor(((double)t->img_sample_sz[0] - (double)aug_sz[0]) / 2);
    g.win_c = -(int)std::floor(((double)t->img_sample_sz[1] - (double)aug_sz[1]) / 2);
    sample_location(t, &g);            // informational; the init sample's position is init_sample_pos / init_sample_scale
    g.sample_pos[0] = init_pos[0]; g.sample_pos[1] = init_pos[1]; g.sample_scale = init_scale;
    if (init_crop) *init_crop = g;
    float box[4];
    iounet_box(t, t->pos, t->target_sz, init_pos, init_scale, box);                                 // init_target_boxes, dimp.py:400-408
    if (init_target_box) std::memcpy(init_target_box, box, sizeof(box));
    // init_memory (dimp.py:410-427) with one init sample
    std::fill(t->sw.begin(), t->sw.end(), 0.f);
    t->sw[0] = 1.0f;
    t->num_init = 1; t->num_stored = 1; t->prev_replace_ind = -1;
    t->initialized = 1;
    return 0;
}

extern "C" int b200trk_dimp_tracker_adopt(b200trk_dimp_tracker_t* t, int H, int W, const float pos[2], const float target_sz[2], float target_scale,
                                          const float base_target_sz[2], float min_scale_factor, float max_scale_factor,
                                          const float* sample_weights, int num_stored, int num_init, int previous_replace_ind, int frame_num) {
    B200_REQUIRE(t && pos && target_sz && base_target_sz && sample_weights, "dimp_tracker_adopt: null pointer");
    B200_REQUIRE(num_init >= 1 && num_stored >= num_init && num_init <= (int)t->sw.size(), "dimp_tracker_adopt: bad sample counts");
    for (int i = 0; i < 2; ++i) { t->pos[i] = pos[i]; t->target_sz[i] = target_sz[i]; t->base_target_sz[i] = base_target_sz[i]; }
    t->image_sz[0] = (float)H; t->image_sz[1] = (float)W;
    t->target_scale = target_scale; t->min_scale_factor = min_scale_factor; t->max_scale_factor = max_scale_factor;
    for (size_t i = 0; i < t->sw.size(); ++i) t->sw[i] = sample_weights[i];
    t->num_stored = num_stored; t->num_init = num_init; t->prev_replace_ind = previous_replace_ind; t->frame_num = frame_num;
    t->initialized = 1;
    return 0;
}

extern "C" int b200trk_dimp_tracker_state(const b200trk_dimp_tracker_t* t, float out[9]) {
    B200_REQUIRE(t && out, "dimp_tracker_state: null pointer");
    out[0] = t->pos[0]; out[1] = t->pos[1]; out[2] = t->target_sz[0]; out[3] = t->target_sz[1]; out[4] = t->target_scale;
    out[5] = t->base_target_sz[0]; out[6] = t->base_target_sz[1]; out[7] = t->min_scale_factor; out[8] = t->max_scale_factor;
    return 0;
}

extern "C" int b200trk_dimp_tracker_plan_crop(b200trk_dimp_tracker_t* t, b200trk_crop_geom_t* g) {
    B200_REQUIRE(t && g, "dimp_tracker_plan_crop: null pointer");
    B200_REQUIRE(t->initialized, "dimp_tracker_plan_crop: tracker not initialised");
    // get_centered_sample_pos (dimp.py:184-188)
    float cpos[2];
    for (int i = 0; i < 2; ++i) {
        const float m = std::fmod(t->feature_sz[i] + t->kernel_size[i], 2.0f);
        cpos[i] = t->pos[i] + ((m * t->target_scale) * t->img_sample_sz[i]) / (2.0f * t->feature_sz[i]);
    }
    const float s = t->target_scale * 1.0f;                                                       // target_scale * scale_factors (ones(1))
    const float sample_sz[2] = {s * t->img_sample_sz[0], s * t->img_sample_sz[1]};
    plan_patch(cpos, sample_sz, t->img_sample_sz, (int)t->image_sz[0], (int)t->image_sz[1], g);
    sample_location(t, g);
    return 0;
}

// DiMP.update_sample_weights (dimp.py:445-484), float32 like the reference's tensor; returns the slot to overwrite
static int update_sample_weights(b200trk_dimp_tracker* t, double lr) {
    std::vector<float>& sw = t->sw;
    const int size = (int)sw.size();
    double init_w = t->p.init_samples_minimum_weight;
    const bool has_init_w = init_w != 0.0;
    const int s_ind = has_init_w ? t->num_init : 0;
    int r_ind;
    auto sum = [&](int a, int b) { float s = 0.f; for (int i = a; i < b; ++i) s += sw[i]; return s; };
    if (t->num_stored == 0 || lr == 1.0) {
        std::fill(sw.begin(), sw.end(), 0.f);
        sw[0] = 1.f; r_ind = 0;
    } else {
        if (t->num_stored < size) {
            r_ind = (int)t->num_stored;
        } else {
            r_ind = s_ind;
            for (int i = s_ind + 1; i < size; ++i) if (sw[i] < sw[r_ind]) r_ind = i;                 // torch.min: first minimum
        }
        const float one_minus = f32(1.0 - lr);
        if (t->prev_replace_ind < 0) {
            for (float& v : sw) v = v / one_minus;
            sw[r_ind] = f32(lr);
        } else {
            sw[r_ind] = sw[t->prev_replace_ind] / one_minus;
        }
    }
    const float tot = sum(0, size);
    for (float& v : sw) v = v / tot;
    if (has_init_w && sum(0, t->num_init) < f32(init_w)) {
        const float d = f32(init_w) + sum(t->num_init, size);
        for (float& v : sw) v = v / d;
        const float iw = f32(init_w / t->num_init);
        for (int i = 0; i < t->num_init; ++i) sw[i] = iw;
    }
    return r_ind;
}

// DiMP.update_state (dimp.py:486-497)
static void update_state(b200trk_dimp_tracker* t, const float new_pos[2], const float* new_scale) {
    if (new_scale) {
        t->target_scale = std::fmin(std::fmax(*new_scale, t->min_scale_factor), t->max_scale_factor);
        for (int i = 0; i < 2; ++i) t->target_sz[i] = t->base_target_sz[i] * t->target_scale;
    }
    const float ir = f32(t->p.target_inside_ratio - 0.5);
    for (int i = 0; i < 2; ++i) {
        const float off = ir * t->target_sz[i];
        t->pos[i] = std::fmax(std::fmin(new_pos[i], t->image_sz[i] - off), off);
    }
}

static bool refines(const b200trk_dimp_tracker* t) { return t->p.use_iou_net != 0; }

// Phase 1 of a frame's host work: everything between localize_target and refine_target_box (dimp.py:97-128)
extern "C" int b200trk_dimp_tracker_commit_localize(b200trk_dimp_tracker_t* t, const b200trk_crop_geom_t* g, const b200trk_loc_result_t* loc,
                                                    b200trk_frame_info_t* info) {
    B200_REQUIRE(t && g && loc && info, "dimp_tracker_commit: null pointer");
    B200_REQUIRE(t->initialized, "dimp_tracker_commit: tracker not initialised");
    t->frame_num += 1;                                                                             // dimp.py:97
    std::memset(info, 0, sizeof(*info));
    info->loc = *loc; info->crop = *g; info->flag = loc->flag; info->max_score = loc->max_score; info->replace_ind = -1;
    const float sample_scale = g->sample_scale;
    // translation (dimp.py:243-257)
    float new_pos[2];
    const int rc[2] = {loc->use_second ? loc->r2 : loc->r1, loc->use_second ? loc->c2 : loc->c1};
    for (int i = 0; i < 2; ++i) {
        const float output_sz = t->score_sz[i] - std::fmod(t->kernel_size[i] + 1.0f, 2.0f);
        const float center = (t->score_sz[i] - 1.0f) / 2.0f;
        const float disp = (float)rc[i] - center;
        const float tv = (disp * (t->img_sample_sz[i] / output_sz)) * sample_scale;
        new_pos[i] = g->sample_pos[i] + tv;
    }
    if (loc->flag != 4) {
        if (refines(t)) update_state(t, new_pos, nullptr);          // the IoUNet sets the size (dimp.py:120-124)
        else update_state(t, new_pos, &sample_scale);               // dimp.py:125-126
    }
    return 0;
}

static float next_uniform(b200trk_dimp_tracker* t, size_t i) {
    if (i < t->noise.size()) return t->noise[i];
    t->rng ^= t->rng << 13; t->rng ^= t->rng >> 7; t->rng ^= t->rng << 17;                        // xorshift64: the tracker's own generator
    return (float)((t->rng >> 40) & 0xFFFFFF) * (1.0f / 16777216.0f);
}

// The proposals of refine_target_box (dimp.py:654-674): the classifier's box followed by num_init_random_boxes jittered copies
extern "C" int b200trk_dimp_tracker_proposals(b200trk_dimp_tracker_t* t, const b200trk_crop_geom_t* g, float* boxes_out, int* count) {
    B200_REQUIRE(t && g && boxes_out && count, "dimp_tracker_proposals: null pointer");
    const int nr = t->p.num_init_random_boxes;
    B200_REQUIRE(nr >= 0 && nr + 1 <= 16, "dimp_tracker_proposals: num_init_random_boxes=%d (at most 15)", nr);
    float ib[4];
    iounet_box(t, t->pos, t->target_sz, g->sample_pos, g->sample_scale, ib);
    for (int i = 0; i < 4; ++i) boxes_out[i] = ib[i];
    if (nr > 0) {
        const float square = std::sqrt(ib[2] * ib[3]);
        const float rf[4] = {square * f32(t->p.box_jitter_pos), square * f32(t->p.box_jitter_pos), square * f32(t->p.box_jitter_sz),
                             square * f32(t->p.box_jitter_sz)};
        const float min_edge = std::fmin(ib[2], ib[3]) / 3.0f;
        for (int r = 0; r < nr; ++r) {
            float rb[4];
            for (int i = 0; i < 4; ++i) rb[i] = (next_uniform(t, (size_t)r * 4 + i) - 0.5f) * rf[i];
            float* o = boxes_out + 4 * (r + 1);
            for (int i = 0; i < 2; ++i) {
                const float nsz = std::fmax(ib[2 + i] + rb[2 + i], min_edge);
                const float nc = (ib[i] + ib[2 + i] / 2.0f) + rb[i];
                o[i] = nc - nsz / 2.0f; o[2 + i] = nsz;
            }
        }
    }
    t->noise.clear();
    *count = nr + 1;
    return 0;
}

// Phase 2: after optimize_boxes -- filter, top-k mean, new position / size / scale (dimp.py:679-721)
extern "C" int b200trk_dimp_tracker_commit_refine(b200trk_dimp_tracker_t* t, const b200trk_crop_geom_t* g, const float* boxes, const float* iou,
                                                  int count, b200trk_frame_info_t* info) {
    B200_REQUIRE(t && g && boxes && iou && info && count >= 1 && count <= 16, "dimp_tracker_commit_refine: bad argument");
    float bx[16][4]; float io[16]; int n = 0;
    const float mar = f32(t->p.maximal_aspect_ratio), imar = f32(1.0 / t->p.maximal_aspect_ratio);
    for (int r = 0; r < count; ++r) {
        const float w = std::fmax(boxes[4 * r + 2], 1.0f), h = std::fmax(boxes[4 * r + 3], 1.0f);    // output_boxes[:, 2:].clamp_(1)
        const float ar = w / h;
        if (ar < mar && ar > imar) { bx[n][0] = boxes[4 * r]; bx[n][1] = boxes[4 * r + 1]; bx[n][2] = w; bx[n][3] = h; io[n] = iou[r]; ++n; }
    }
    if (n == 0) return 0;                                                                               // "If no box found"
    const int k = std::min(t->p.iounet_k > 0 ? t->p.iounet_k : 5, n);
    bool used[16] = {false};
    float pb[4] = {0, 0, 0, 0}, piou = 0.f;
    for (int j = 0; j < k; ++j) {                          // torch.topk: largest first
        int best = -1;
        for (int r = 0; r < n; ++r) if (!used[r] && (best < 0 || io[r] > io[best])) best = r;
        used[best] = true;
        for (int i = 0; i < 4; ++i) pb[i] += bx[best][i];
        piou += io[best];
    }
    for (int i = 0; i < 4; ++i) pb[i] = pb[i] / (float)k;
    info->predicted_iou = piou / (float)k; info->refined = 1;
    // new position and size (dimp.py:703-721); predicted_box = (x, y, w, h), tracker vectors are (row, col)
    const float cx = pb[0] + pb[2] / 2.0f, cy = pb[1] + pb[3] / 2.0f;
    const float c[2] = {cy, cx}, sz[2] = {pb[3], pb[2]};
    float new_pos[2], new_sz[2];
    for (int i = 0; i < 2; ++i) {
        new_pos[i] = (c[i] - (t->img_sample_sz[i] - 1.0f) / 2.0f) * g->sample_scale + g->sample_pos[i];
        new_sz[i] = sz[i] * g->sample_scale;
    }
    const float new_scale = std::sqrt((new_sz[0] * new_sz[1]) / (t->base_target_sz[0] * t->base_target_sz[1]));
    t->pos_iounet[0] = new_pos[0]; t->pos_iounet[1] = new_pos[1]; t->has_pos_iounet = 1;
    if (t->p.use_iounet_pos_for_learning) { t->pos[0] = new_pos[0]; t->pos[1] = new_pos[1]; }
    t->target_sz[0] = new_sz[0]; t->target_sz[1] = new_sz[1];
    const bool update_scale = t->p.update_scale_when_uncertain || info->flag != 3;
    if (update_scale) t->target_scale = new_scale;
    return 0;
}

// Phase 3: the update half of the frame and the output box (dimp.py:131-175, 605-625)
extern "C" int b200trk_dimp_tracker_commit_update(b200trk_dimp_tracker_t* t, const b200trk_crop_geom_t* g, b200trk_frame_info_t* info,
                                                  float* sample_weights_out) {
    B200_REQUIRE(t && g && info, "dimp_tracker_commit: null pointer");
    const float sample_scale = g->sample_scale;
    const bool not_found = info->flag == 4, uncertain = info->flag == 3, hard_negative = info->flag == 2;
    const bool update_flag = !not_found && !uncertain;
    if (update_flag && t->p.update_classifier) {
        iounet_box(t, t->pos, t->target_sz, g->sample_pos, sample_scale, info->target_box);
        const bool hn_flag = hard_negative && t->p.hard_negative_learning_rate >= 0.0;
        const double lr = hn_flag ? t->p.hard_negative_learning_rate : t->p.learning_rate;
        const int interval = t->p.train_sample_interval > 0 ? t->p.train_sample_interval : 1;
        if (hn_flag || t->frame_num % interval == 0) {
            info->replace_ind = update_sample_weights(t, lr);
            t->prev_replace_ind = info->replace_ind;
            t->num_stored += 1;
            info->updated = 1;
        }
        int num_iter = 0;
        if (hn_flag) num_iter = t->p.net_opt_hn_iter;
        else if ((t->frame_num - 1) % (t->p.train_skipping > 0 ? t->p.train_skipping : 1) == 0) num_iter = t->p.net_opt_update_iter;
        info->num_iter = num_iter > 0 ? num_iter : 0;
        info->learning_rate = f32(lr);
    }
    info->n_stored = (int)std::min<long>(t->num_stored, (long)t->sw.size());
    // "Set the pos of the tracker to iounet pos" (dimp.py:148-150)
    if (refines(t) && !not_found && t->has_pos_iounet) { t->pos[0] = t->pos_iounet[0]; t->pos[1] = t->pos_iounet[1]; }
    // output box (dimp.py:163-171)
    if (t->p.output_not_found_box && not_found) {
        for (int i = 0; i < 4; ++i) info->bbox[i] = -1.f;
    } else {
        info->bbox[0] = t->pos[1] - (t->target_sz[1] - 1.0f) / 2.0f;
        info->bbox[1] = t->pos[0] - (t->target_sz[0] - 1.0f) / 2.0f;
        info->bbox[2] = t->target_sz[1];
        info->bbox[3] = t->target_sz[0];
    }
    if (sample_weights_out) std::memcpy(sample_weights_out, t->sw.data(), t->sw.size() * sizeof(float));
    return 0;
}

// The whole host half of a frame without IoUNet refinement (use_iou_net = False): phases 1 + 3
extern "C" int b200trk_dimp_tracker_commit(b200trk_dimp_tracker_t* t, const b200trk_crop_geom_t* g, const b200trk_loc_result_t* loc,
                                           b200trk_frame_info_t* info, float* sample_weights_out) {
    B200_REQUIRE(t && !refines(t), "dimp_tracker_commit: use_iou_net is set -- call commit_localize / proposals / commit_refine / commit_update");
    if (int e = b200trk_dimp_tracker_commit_localize(t, g, loc, info)) return e;
    return b200trk_dimp_tracker_commit_update(t, g, info, sample_weights_out);
}

extern "C" int b200trk_dimp_tracker_set_proposal_noise(b200trk_dimp_tracker_t* t, const float* u01, int count) {
    B200_REQUIRE(t && u01 && count >= 0 && count <= 64, "dimp_tracker_set_proposal_noise: bad argument");
    t->noise.assign(u01, u01 + count);
    return 0;
}

extern "C" int b200trk_dimp_tracker_attach_iounet(b200trk_dimp_tracker_t* t, b200trk_iou_predictor_t* pred, const float* modulation3,
                                                  const float* modulation4) {
    B200_REQUIRE(t && pred && modulation3 && modulation4, "dimp_tracker_attach_iounet: null pointer");
    B200_REQUIRE(t->st, "dimp_tracker_attach_iounet: host-logic-only tracker");
    B200_REQUIRE(!t->iou_pred, "dimp_tracker_attach_iounet: already attached");
    if (int e = b200trk_net_iou_dims(t->st->net, t->iou_dims)) return e;
    B200_REQUIRE(t->iou_dims[0] > 0, "dimp_tracker_attach_iounet: the network has no IoU feature branch (b200trk_net_attach_iou_head)");
    const int C3 = t->iou_dims[0], C4 = t->iou_dims[3];
    B200_CHECK_CUDA(cudaMalloc((void**)&t->mod3, C3 * sizeof(float)));
    B200_CHECK_CUDA(cudaMalloc((void**)&t->mod4, C4 * sizeof(float)));
    B200_CHECK_CUDA(cudaMemcpy(t->mod3, modulation3, C3 * sizeof(float), cudaMemcpyHostToDevice));
    B200_CHECK_CUDA(cudaMemcpy(t->mod4, modulation4, C4 * sizeof(float), cudaMemcpyHostToDevice));
    B200_CHECK_CUDA(cudaMalloc((void**)&t->iou3, (size_t)C3 * t->iou_dims[1] * t->iou_dims[2] * sizeof(float)));
    B200_CHECK_CUDA(cudaMalloc((void**)&t->iou4, (size_t)C4 * t->iou_dims[4] * t->iou_dims[5] * sizeof(float)));
    B200_CHECK_CUDA(cudaMalloc((void**)&t->boxes_dev, 16 * 5 * sizeof(float)));
    B200_CHECK_CUDA(cudaMallocHost((void**)&t->boxes_host, 16 * 5 * sizeof(float)));
    t->iou_pred = pred;
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------------------
// device launches
// ---------------------------------------------------------------------------------------------------------------------------
extern "C" int b200trk_sample_patch(const uint8_t* image_dev, int H, int W, const b200trk_crop_geom_t* g, int win_h, int win_w, float* out,
                                    b200trk_stream_t stream) {
    B200_REQUIRE(image_dev && g && out, "sample_patch: null pointer");
    B200_REQUIRE(H > 0 && W > 0 && win_h > 0 && win_w > 0 && g->df >= 1 && g->in_h >= 1 && g->in_w >= 1 && g->out_h >= 1 && g->out_w >= 1,
                 "sample_patch: bad geometry");
    B200_REQUIRE(g->win_r >= 0 && g->win_c >= 0 && g->win_r + win_h <= g->out_h && g->win_c + win_w <= g->out_w,
                 "sample_patch: window [%d+%d, %d+%d] outside the %dx%d resampled patch", g->win_r, win_h, g->win_c, win_w, g->out_h, g->out_w);
    dim3 grid((win_w + 31) / 32, (win_h + 7) / 8);
    sample_patch_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(image_dev, H, W, *g, win_h, win_w, out);
    B200_LAUNCH_CHECK();
    return 0;
}

extern "C" int b200trk_dimp_localize(const float* scores, int S, int Ho, int Wo, const b200trk_dimp_params_t* p, const float* neigh,
                                     const float* prev_vec, b200trk_loc_result_t* result_dev, b200trk_stream_t stream) {
    B200_REQUIRE(scores && p && result_dev, "dimp_localize: null pointer");
    B200_REQUIRE(S >= 1 && S <= 8 && Ho > 0 && Wo > 0, "dimp_localize: S=%d (1..8), map %dx%d", S, Ho, Wo);
    B200_REQUIRE(!p->advanced_localization || (neigh && prev_vec), "dimp_localize: advanced localisation needs neigh / prev_vec");
    const LocArgs a = make_loc_args(S, Ho, Wo, p, neigh, prev_vec);
    localize_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(scores, a, result_dev);
    B200_LAUNCH_CHECK();
    return 0;
}

static int upload_image(b200trk_dimp_tracker* t, const uint8_t* image, int H, int W, cudaStream_t st) {
    const size_t bytes = (size_t)H * W * 3;
    if (bytes > t->img_cap) {
        if (t->img_dev) B200_CHECK_CUDA(cudaFree(t->img_dev));
        t->img_dev = nullptr; t->img_cap = 0;
        B200_CHECK_CUDA(cudaMalloc((void**)&t->img_dev, bytes));
        t->img_cap = bytes;
    }
    // The previous frame's steepest-descent update may still be running on `st` (track returns as soon as the box is known); the
    // frame travels on a copy stream of its own meanwhile. (The device image is free: the previous crop kernel finished before the
    // previous call returned, which synchronised on its localisation result.)
    if (!t->copy_stream) {
        B200_CHECK_CUDA(cudaStreamCreateWithFlags(&t->copy_stream, cudaStreamNonBlocking));
        B200_CHECK_CUDA(cudaEventCreateWithFlags(&t->ev_img, cudaEventDisableTiming));
    }
    B200_CHECK_CUDA(cudaMemcpyAsync(t->img_dev, image, bytes, cudaMemcpyHostToDevice, t->copy_stream));
    B200_CHECK_CUDA(cudaEventRecord(t->ev_img, t->copy_stream));
    B200_CHECK_CUDA(cudaStreamWaitEvent(st, t->ev_img, 0));
    return 0;
}

extern "C" int b200trk_dimp_tracker_initialize_host(b200trk_dimp_tracker_t* t, const uint8_t* image, int H, int W, const double init_bbox[4],
                                                    b200trk_stream_t stream) {
    B200_REQUIRE(t && image && init_bbox, "dimp_tracker_initialize_host: null pointer");
    B200_REQUIRE(t->st, "dimp_tracker_initialize_host: host-logic-only tracker (created without a device state)");
    cudaStream_t st = (cudaStream_t)stream;
    b200trk_dimp_state* s = t->st;
    b200trk_crop_geom_t g; float box[4];
    if (int e = b200trk_dimp_tracker_init_state(t, H, W, init_bbox, &g, box)) return e;
    if (int e = upload_image(t, image, H, W, st)) return e;
    const int iss = t->p.image_sample_size;
    if (int e = b200trk_sample_patch(t->img_dev, H, W, &g, iss, iss, s->crop, stream)) return e;
    if (int e = b200trk_net_forward(s->net, s->crop, 1, nullptr, nullptr, s->clf, stream)) return e;
    // FilterInitializerZero (dimp.py:601-602) + net_opt_iter iterations on the single init sample (sample_weight = None -> 1/n)
    B200_CHECK_CUDA(cudaMemsetAsync(s->filter, 0, (size_t)s->Cc * s->ksz * s->ksz * sizeof(float), st));
    B200_CHECK_CUDA(cudaMemsetAsync(s->sw, 0, (size_t)s->memory_size * sizeof(float), st));
    const float one = 1.f;
    if (int e = dimp_state_update(s, 0, 0, box, &one, 1, t->p.net_opt_iter > 0 ? t->p.net_opt_iter : 0, st)) return e;
    B200_CHECK_CUDA(cudaStreamSynchronize(st));
    return 0;
}

static int track_frame(b200trk_dimp_tracker* t, const uint8_t* image, bool on_device, int H, int W, b200trk_frame_info_t* info,
                       b200trk_stream_t stream) {
    B200_REQUIRE(t && image && info, "dimp_track_host: null pointer");
    B200_REQUIRE(t->st, "dimp_track_host: host-logic-only tracker (created without a device state)");
    B200_REQUIRE(t->initialized, "dimp_track_host: tracker not initialised");
    B200_REQUIRE((float)H == t->image_sz[0] && (float)W == t->image_sz[1], "dimp_track_host: frame is %dx%d, the sequence started with %dx%d",
                 H, W, (int)t->image_sz[0], (int)t->image_sz[1]);
    cudaStream_t st = (cudaStream_t)stream;
    b200trk_dimp_state* s = t->st;
    b200trk_crop_geom_t g;
    if (int e = b200trk_dimp_tracker_plan_crop(t, &g)) return e;
    if (!on_device) { if (int e = upload_image(t, image, H, W, st)) return e; }
    const int iss = t->p.image_sample_size;
    if (int e = b200trk_sample_patch(on_device ? image : t->img_dev, H, W, &g, iss, iss, s->crop, stream)) return e;
    if (int e = b200trk_net_forward(s->net, s->crop, 1, nullptr, nullptr, s->clf, stream)) return e;
    if (int e = b200trk_apply_filter(s->clf, s->filter, s->scores, 1, s->Cc, s->Hc, s->Wc, s->ksz, nullptr, nullptr, stream)) return e;
    // target_neigh_sz (dimp.py:265) and prev_target_vec (dimp.py:283) of the single scale
    float neigh[2], pv[2];
    for (int i = 0; i < 2; ++i) {
        const float output_sz = t->score_sz[i] - std::fmod(t->kernel_size[i] + 1.0f, 2.0f);
        neigh[i] = (f32(t->p.target_neighborhood_scale) * (t->target_sz[i] / g.sample_scale)) * (output_sz / t->img_sample_sz[i]);
        pv[i] = (t->pos[i] - g.sample_pos[i]) / ((t->img_sample_sz[i] / output_sz) * g.sample_scale);
    }
    if (int e = b200trk_dimp_localize(s->scores, 1, s->Ho, s->Wo, &t->p, neigh, pv, t->loc_dev, stream)) return e;
    B200_CHECK_CUDA(cudaMemcpyAsync(t->loc_host, t->loc_dev, sizeof(b200trk_loc_result_t), cudaMemcpyDeviceToHost, st));
    B200_CHECK_CUDA(cudaStreamSynchronize(st));
    if (int e = b200trk_dimp_tracker_commit_localize(t, &g, t->loc_host, info)) return e;
    if (refines(t) && info->flag != 4) {
        // refine_target_box (dimp.py:650-723): IoU features of this crop, proposals up, box optimisation on the device, 320 bytes back
        B200_REQUIRE(t->iou_pred, "dimp_track_host: use_iou_net is set but no IoU predictor is attached (b200trk_dimp_tracker_attach_iounet)");
        int R = 0;
        if (int e = b200trk_dimp_tracker_proposals(t, &g, t->boxes_host, &R)) return e;
        B200_CHECK_CUDA(cudaMemcpyAsync(t->boxes_dev, t->boxes_host, (size_t)R * 4 * sizeof(float), cudaMemcpyHostToDevice, st));
        if (int e = b200trk_net_iou_from_arena(s->net, 1, t->iou3, t->iou4, stream)) return e;
        if (int e = b200trk_iou_refine(t->iou_pred, t->mod3, t->mod4, t->iou3, t->iou_dims[1], t->iou_dims[2], t->iou4, t->iou_dims[4],
                                       t->iou_dims[5], t->boxes_dev, R, t->p.box_refinement_iter, f32(t->p.box_refinement_step_length),
                                       f32(t->p.box_refinement_step_decay), t->p.box_refinement_relative, t->boxes_dev + 64, stream)) return e;
        B200_CHECK_CUDA(cudaMemcpyAsync(t->boxes_host, t->boxes_dev, 80 * sizeof(float), cudaMemcpyDeviceToHost, st));
        B200_CHECK_CUDA(cudaStreamSynchronize(st));
        if (int e = b200trk_dimp_tracker_commit_refine(t, &g, t->boxes_host, t->boxes_host + 64, R, info)) return e;
    }
    if (int e = b200trk_dimp_tracker_commit_update(t, &g, info, nullptr)) return e;
    if (info->updated) {
        if (int e = dimp_state_update(s, 0, info->replace_ind, info->target_box, t->sw.data(), info->n_stored, info->num_iter, st)) return e;
    } else if (info->num_iter > 0) {
        // update_classifier can optimise without storing a sample (train_sample_interval > 1): run the iterations only
        if (int e = dimp_sd_gn_pitched(s->filter, s->filter, s->memory, s->mem_pitch, s->boxes, s->sw, info->n_stored, s->Cc, s->Hc, s->Wc,
                                       s->ksz, info->num_iter, s->luts, s->luts + s->num_bins, s->luts + 2 * s->num_bins, s->num_bins,
                                       s->bin_displacement, s->feat_stride, s->step_length, s->reg_weight, s->alpha_eps, nullptr, nullptr, st))
            return e;
    }
    return 0;
}

extern "C" int b200trk_dimp_track_host(b200trk_dimp_tracker_t* t, const uint8_t* image, int H, int W, b200trk_frame_info_t* info,
                                       b200trk_stream_t stream) {
    return track_frame(t, image, false, H, W, info, stream);
}

extern "C" int b200trk_dimp_track_device(b200trk_dimp_tracker_t* t, const uint8_t* image_dev, int H, int W, b200trk_frame_info_t* info,
                                         b200trk_stream_t stream) {
    return track_frame(t, image_dev, true, H, W, info, stream);
}

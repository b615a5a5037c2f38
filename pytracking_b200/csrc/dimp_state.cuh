// Per-sequence device state of the DiMP online model (shared by dimp_state.cu and dimp_tracker.cu).
#pragma once
#include "net.cuh"
#include "sd_common.cuh"
#include <vector>

struct b200trk_dimp_state {
    b200trk_net* net = nullptr;
    int memory_size = 0, ksz = 4, Cc = 0, Hc = 0, Wc = 0, Ho = 0, Wo = 0, num_bins = 0, max_batch = 1;
    int mem_pitch = 0;      // floats between two channel planes of the sample memory: H*W rounded up to 32, so that every 32-pixel
                            // TMA row of the optimiser is one 128-byte L2 line instead of straddling two
    float bin_displacement = 0.1f, feat_stride = 16.f, step_length = 1.f, reg_weight = 0.01f, alpha_eps = 0.f;
    float *filter = nullptr, *memory = nullptr, *boxes = nullptr, *sw = nullptr, *clf = nullptr, *scores = nullptr;
    float *crop = nullptr, *maxval = nullptr, *luts = nullptr;
    int64_t* maxidx = nullptr;
    std::vector<void*> owned;
    // pinned staging ring for the small per-update H2D payload (box + sample weights): a slot is rewritten only after the copy
    // that read it has completed (its event), so back-to-back asynchronous update calls never race with their own uploads
    static constexpr int NSTAGE = 4;
    float* stage[NSTAGE] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t stage_ev[NSTAGE] = {nullptr, nullptr, nullptr, nullptr};
    int stage_next = 0;
};

namespace b200trk {
int dimp_state_update(b200trk_dimp_state* s, int scale_ind, int replace_ind, const float* target_box_host,
                      const float* sample_weights_host, int n_stored, int num_iter, cudaStream_t st);
}

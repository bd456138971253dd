// Persistent cooperative "step kernel": the whole Pi-0 control step as ONE launch.
//
// At batch 1 the step is a chain of ~600 small dependent operations; as separate kernels each costs
// ~10 us of launch / ramp / drain latency (measured: 6.3 ms for 625 kernels whose useful work is ~1 ms).
// Here one CTA per SM stays resident (TMEM and the shared-memory TMA ring allocated once) and walks
// a device-resident op list; the work items of an op are dealt round-robin to the CTAs, and a
// grid-wide barrier (one atomic + acquire spin per CTA, ~1-2 us) replaces the kernel boundary.
// Independent ops (VLM and proprio streams of the same layer) share a barrier.
//
// The op bodies are the same device functions the stand-alone kernels run (bodies.cuh,
// gemm_body.cuh), so results are bit-identical to the multi-kernel path.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "gemm_tc.h"
#include "kernels.h"

namespace blurr {

enum StepOpType {
    OP_GEMM = 0, OP_CONSUMER, OP_BIAS_ACT, OP_ROPE_KV, OP_ATTN_SIGLIP, OP_ATTN_PREFILL, OP_ATTN_FEWQ,
    OP_EMBED_MERGE, OP_SMALL_K, OP_ACTION_TAIL, OP_CLAMP,
};

struct BiasActArgs { const float* partial; int splitk, T, N, ldp; const bf16* bias; int act; float scale; bf16* out; int ldo; };
struct EmbedMergeArgs {
    const int64_t* ids; int seq; const bf16* table; long long vocab; const bf16* img; int n_img, hidden;
    long long image_token, pad_token; float inv_div, normalizer; bf16* out; int* err_flag;
};
struct SmallKArgs {
    const bf16* x; int T, K; const bf16* W; const bf16* bias; int N; float scale; bf16* y; int ldy, col_off;
    const bf16* time_row; int time_cols;
};
struct ActionTailArgs { const bf16* xn; int T, hidden; const bf16* W; const bf16* bias; int action_dim; float dt; bf16* action; bf16* vel_tap; };
struct ClampArgs { const bf16* src; bf16* dst; int n, do_clamp; float clip; };

union StepOpArgs {
    GemmDev gemm;
    ConsumerArgs consumer;
    BiasActArgs bias_act;
    RopeKvArgs rope;
    AttnMmaArgs attn;
    JointAttnArgs fewq;
    EmbedMergeArgs embed;
    SmallKArgs small_k;
    ActionTailArgs tail;
    ClampArgs clamp;
};

struct StepOpHot {            // copied to shared memory at the start of every op
    int type, epi, gx, gy, gz, barrier_after, pad0, pad1;
    StepOpArgs u;
};

struct alignas(128) StepOp {
    CUtensorMap tmap_w;       // OP_GEMM only; TMA reads the descriptors straight from global memory
    CUtensorMap tmap_x;
    StepOpHot hot;
};

struct StepProgram {
    std::vector<StepOp> ops;
    StepOp* d_ops = nullptr;
    size_t d_capacity = 0;
    unsigned* d_sync = nullptr;     // [0] barrier counter, [1] error flag
    int n_barriers = 0;
};

// Upload the op list (synchronous) — called once per (batch, steps) schedule.
int step_program_upload(StepProgram& prog, std::string* err);
void step_program_free(StepProgram& prog);
// Reset the barrier counter and launch the cooperative kernel on `stream`.
int step_program_launch(const StepProgram& prog, cudaStream_t stream, std::string* err);
// 0 = ok; non-zero: a grid barrier or a GEMM pipeline wait expired inside the last launches (synchronises).
int step_program_take_error(const StepProgram& prog);

}  // namespace blurr

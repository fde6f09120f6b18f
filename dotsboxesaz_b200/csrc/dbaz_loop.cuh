// dbaz_loop.cuh -- one CUDA graph per search: the adaptive wave loop without the host.
//
//   begin kernel (initial rung from the busy-tree count)
//   WHILE node (condition: trees still busy) {
//       SWITCH node over the batch ladder: body r = the captured graph of `waves(r)` [step -> evaluator] waves at rung r
//       decide kernel: reads the wave counters the last step published, picks the next rung (the policy of
//                      Engine._pick_rows: most rows served per microsecond of measured evaluator time, a rung up to 10 %
//                      short when clearly cheaper per row), sets both condition values
//   }
//
// The rung graphs are captured by the host (PyTorch stream capture of the engine's step kernel and the evaluator's
// kernels) and cloned into the switch bodies as child graphs.  The host launches ONE graph per UCT_search of all trees and
// never waits inside it (engine.py waited for an event per replay before); the decision uses the counters of the replay
// that just finished instead of the one before it.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dbaz {

constexpr int LOOP_MAX_RUNGS = 48;

struct LoopCtl {
    int n_rungs;
    int rows[LOOP_MAX_RUNGS];        // evaluator batch of rung r, DESCENDING
    float us[LOOP_MAX_RUNGS];        // measured evaluator time of rung r (microseconds)
    float row_margin, undersize, undersize_gain, wave_overhead_us;
    int max_iters;                   // safety bound on the replays of one launch (a loop must never hang the GPU)
    // state
    int iters;
    unsigned int replays[LOOP_MAX_RUNGS];  // replays of every rung since the counters were last read by the host
};

__host__ __device__ __forceinline__ int loop_pick(const LoopCtl* c, int want) {
    // the smallest rung that holds `want` rows ... (rows are descending: the last index with rows >= want)
    int fit_lo = 0;  // rungs [0, fit_hi] hold the rows
    int fit_hi = -1;
    for (int r = 0; r < c->n_rungs; ++r) if (c->rows[r] >= want) fit_hi = r;
    if (fit_hi < 0) return 0;
    // ... or, among the rungs that hold them, the one that serves the most rows per microsecond
    float best_sc = -1.0f;
    int best = fit_hi;
    for (int r = fit_lo; r <= fit_hi; ++r) {
        const float sc = (float)(c->rows[r] < want ? c->rows[r] : want) / (c->us[r] + c->wave_overhead_us);
        if (sc >= best_sc) { best_sc = sc; best = r; }  // ties: the smaller rung (later index)
    }
    // a rung up to 10 % short of the rows wanted, if clearly cheaper per row served (its surplus leaves wait a wave)
    for (int r = fit_hi + 1; r < c->n_rungs; ++r) {
        if ((float)c->rows[r] >= c->undersize * (float)want && c->rows[r] < want) {
            const float sc = (float)c->rows[r] / (c->us[r] + c->wave_overhead_us);
            if (sc > best_sc * c->undersize_gain) { best_sc = sc / c->undersize_gain * 1.0001f; best = r; }
        }
    }
    return best;
}

__global__ void k_loop_begin(int* ctr, LoopCtl* c, cudaGraphConditionalHandle h_while, cudaGraphConditionalHandle h_switch) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int busy = ctr[5];  // trees that take part in this search (k_search_begin counted them)
    c->iters = 0;
    ctr[6] = 0;
    int r = 0;
    for (int i = 0; i < c->n_rungs; ++i) if (c->rows[i] >= busy) r = i;
    cudaGraphSetConditional(h_switch, (unsigned)r);
    cudaGraphSetConditional(h_while, busy > 0 ? 1u : 0u);
    if (busy > 0) atomicAdd(&c->replays[r], 1u);
}

__global__ void k_loop_decide(int* ctr, LoopCtl* c, cudaGraphConditionalHandle h_while, cudaGraphConditionalHandle h_switch) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int busy = ctr[5];
    const int rows_max = ctr[6];  // the largest number of rows a wave of the replay asked for
    ctr[6] = 0;
    c->iters += 1;
    const bool go = busy > 0 && c->iters < c->max_iters;
    int r = 0;
    if (go) {
        int want = (int)((float)rows_max * c->row_margin) + 32;
        if (want > busy) want = busy;
        r = loop_pick(c, want);
        atomicAdd(&c->replays[r], 1u);
    }
    cudaGraphSetConditional(h_switch, (unsigned)r);
    cudaGraphSetConditional(h_while, go ? 1u : 0u);
}

}  // namespace dbaz

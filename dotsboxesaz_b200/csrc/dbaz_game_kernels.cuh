// dbaz_game_kernels.cuh -- batched Dots & Boxes rules, one thread per game, states resident
// in HBM as 32-byte packed records.  HBM-bound byte/bit work: coalesced 32-byte sector loads
// and stores, no shared memory needed.
#pragma once
#include "dbaz_device.cuh"

namespace dbaz {

__global__ void k_game_init(Board b, dbaz_state* __restrict__ states, int64_t n) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    dbaz_state s;
    state_init(b, s);
    states[i] = s;
}

// get_valid_moves: uint8[n][A]; each thread writes one game's row (A bytes)
template <int NW>
__global__ void k_game_valid(Board b, const dbaz_state* __restrict__ states, uint8_t* __restrict__ out, int64_t n) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    dbaz_state s = states[i];
    for (int a = 0; a < b.A; ++a) out[i * b.A + a] = state_legal<NW>(b, s, a) ? 1 : 0;
}

template <int NW>
__global__ void k_game_play(Board b, dbaz_state* __restrict__ states, const int32_t* __restrict__ moves,
                            int32_t* __restrict__ n_closed, int32_t* __restrict__ closed_lc, int64_t n) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    dbaz_state s = states[i];
    int a = moves[i];
    int lc[2][2] = {{-1, -1}, {-1, -1}};
    int nc = -1;
    if (state_legal<NW>(b, s, a)) {
        Mask<NW> box[2];
        action_boxes<NW>(b, a, box, lc);
        Mask<NW> e = load_edges<NW>(s);
        mask_set(e, a);
        bool c0 = mask_any(box[0]) && mask_covers(e, box[0]);
        bool c1 = mask_any(box[1]) && mask_covers(e, box[1]);
        nc = (int)c0 + (int)c1;
        if (!c0) { lc[0][0] = lc[1][0]; lc[0][1] = lc[1][1]; lc[1][0] = lc[1][1] = -1; if (!c1) lc[0][0] = lc[0][1] = -1; }
        else if (!c1) lc[1][0] = lc[1][1] = -1;
        state_apply<NW>(s, a, nc);
        states[i] = s;
    }
    n_closed[i] = nc;
    if (closed_lc) {
        if (nc < 0) lc[0][0] = lc[0][1] = lc[1][0] = lc[1][1] = -1;
        reinterpret_cast<int4*>(closed_lc)[i] = make_int4(lc[0][0], lc[0][1], lc[1][0], lc[1][1]);
    }
}

__global__ void k_game_result(const dbaz_state* __restrict__ states, int8_t* __restrict__ out, int64_t n) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    dbaz_state s = states[i];
    out[i] = (int8_t)state_result(s);
}

// get_features for n states: one warp per state so the row is written with coalesced vector stores
template <int NW>
__global__ void k_game_features(Board b, const dbaz_state* __restrict__ states, void* __restrict__ planes, int dtype,
                                int layout, int64_t n) {
    int64_t w = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (w >= n) return;
    dbaz_state s = states[w];
    write_planes_warp<NW>(b, load_edges<NW>(s), (int)(int8_t)s.btc2[s.to_play], planes, w, dtype, layout, lane);
}

// Uniform random legal playout to terminal; the whole game runs in registers, HBM sees the
// state once in and once out (plus the optional move list).
template <int NW>
__global__ void k_game_rollout(Board b, dbaz_state* __restrict__ states, uint64_t seed, uint64_t game0,
                               int32_t* __restrict__ n_plies, uint8_t* __restrict__ moves, int max_plies, int64_t n) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    dbaz_state s = states[i];
    int ply = 0;
    while (state_result(s) == DBAZ_RESULT_NONE) {
        uint64_t l0 = b.real[0] & ~s.edges[0];
        uint64_t l1 = NW == 2 ? (b.real[1] & ~s.edges[1]) : 0ull;
        int k0 = __popcll(l0), k = k0 + __popcll(l1);
        if (k == 0) break;
        uint32_t u = philox_u32(seed, game0 + (uint64_t)i, (uint32_t)ply);
        int pick = (int)__umulhi(u, (uint32_t)k);
        int a = pick < k0 ? nth_set_bit(l0, pick) : 64 + nth_set_bit(l1, pick - k0);
        Mask<NW> box[2];
        int lc[2][2];
        action_boxes<NW>(b, a, box, lc);
        Mask<NW> e = load_edges<NW>(s);
        mask_set(e, a);
        state_apply<NW>(s, a, closed_count<NW>(e, box));
        if (moves && ply < max_plies) moves[i * max_plies + ply] = (uint8_t)a;
        ++ply;
    }
    states[i] = s;
    n_plies[i] = ply;
}

// Deterministic stand-in for the policy/value net (test/bench utility; SURVEY.md 8a KAT):
//   h = hash[0] & 0xffffffff (hash[0] = sum of 1<<move == the low edge word)
//   kind 0: raw_i = float32((h*2654435761 + i*40503) mod 1024) + 1; p = raw / sum(raw);
//           v = float32(((h mod 2001) - 1000) / 1000)
//   kind 1: p_i = 1/A; v = float32((((h*31) mod 5) - 2) / 2)
// One warp per leaf; lanes cover the actions so the prior row is a coalesced store.
__global__ void k_fake_nn(Board b, const dbaz_state* __restrict__ leaves, float* __restrict__ priors,
                          float* __restrict__ values, int kind, int64_t n) {
    int64_t w = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (w >= n) return;
    uint64_t h = leaves[w].edges[0] & 0xffffffffull;
    if (kind == 0) {
        float sum = 0.0f;
        for (int i = 0; i < b.A; ++i) sum += (float)((h * 2654435761ull + (uint64_t)i * 40503ull) % 1024ull) + 1.0f;  // exact
        for (int a = lane; a < b.A; a += 32) {
            float raw = (float)((h * 2654435761ull + (uint64_t)a * 40503ull) % 1024ull) + 1.0f;
            priors[w * b.A + a] = __fdiv_rn(raw, sum);
        }
        if (lane == 0) values[w] = (float)__ddiv_rn((double)(h % 2001ull) - 1000.0, 1000.0);
    } else {
        float u = __fdiv_rn(1.0f, (float)b.A);
        for (int a = lane; a < b.A; a += 32) priors[w * b.A + a] = u;
        if (lane == 0) values[w] = (float)__ddiv_rn((double)((h * 31ull) % 5ull) - 2.0, 2.0);
    }
}

}  // namespace dbaz

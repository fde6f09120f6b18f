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
// state once in and once out (plus the optional move list).  The state is unpacked into scalars: indexing
// boxes_to_close by the player to move would put the whole struct into local memory (80 LDL/STL in the first version of
// this kernel, which is what it then waited for); the two boxes an edge borders come from the engine's action table
// (k_build_act_tab: two 16-byte loads, L1-resident) instead of being rebuilt from the geometry at every ply.
template <int NW>
__global__ void k_game_rollout(Board b, const uint4* __restrict__ act_tab, dbaz_state* __restrict__ states, uint64_t seed, uint64_t game0,
                               int32_t* __restrict__ n_plies, uint8_t* __restrict__ moves, int max_plies, int64_t n) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    dbaz_state s = states[i];
    uint64_t e0 = s.edges[0], e1 = NW == 2 ? s.edges[1] : 0ull;
    int btc0 = s.btc2[0], btc1 = s.btc2[1];   // 2 * boxes_to_close of player 0 / 1
    int tp = s.to_play, jp = s.just_played;
    const uint64_t real0 = b.real[0], real1 = NW == 2 ? b.real[1] : 0ull;
    int ply = 0;
    // get_result() is None: not both zero, and neither player below zero (dots_boxes_game.py:51-59)
    while ((btc0 | btc1) != 0 && btc0 >= 0 && btc1 >= 0) {
        const uint64_t l0 = real0 & ~e0, l1 = real1 & ~e1;
        const int k0 = __popcll(l0), k = k0 + __popcll(l1);
        if (k == 0) break;
        const uint32_t u = philox_u32(seed, game0 + (uint64_t)i, (uint32_t)ply);
        const int pick = (int)__umulhi(u, (uint32_t)k);
        const bool low = pick < k0;
        const int a = (low ? 0 : 64) + nth_set_bit(low ? l0 : l1, low ? pick : pick - k0);
        if (NW == 1 || a < 64) e0 |= 1ull << (a & 63); else e1 |= 1ull << (a & 63);
        const uint4 m0 = act_tab[2 * a], m1 = act_tab[2 * a + 1];
        const uint64_t b00 = (uint64_t)m0.x | ((uint64_t)m0.y << 32), b01 = (uint64_t)m0.z | ((uint64_t)m0.w << 32);
        const uint64_t b10 = (uint64_t)m1.x | ((uint64_t)m1.y << 32), b11 = (uint64_t)m1.z | ((uint64_t)m1.w << 32);
        int nc = 0;  // boxes the edge closes: a box mask is empty where the edge has no box on that side
        nc += ((b00 | b01) != 0 && (e0 & b00) == b00 && (NW == 1 || (e1 & b01) == b01)) ? 1 : 0;
        nc += ((b10 | b11) != 0 && (e0 & b10) == b10 && (NW == 1 || (e1 & b11) == b11)) ? 1 : 0;
        jp = tp;
        if (nc == 0) tp = 1 - tp;
        else if (tp) btc1 -= 2 * nc; else btc0 -= 2 * nc;
        if (moves && ply < max_plies) moves[i * max_plies + ply] = (uint8_t)a;
        ++ply;
    }
    s.edges[0] = e0;
    if (NW == 2) s.edges[1] = e1;
    s.btc2[0] = (int16_t)btc0; s.btc2[1] = (int16_t)btc1;
    s.to_play = (uint8_t)tp; s.just_played = (int8_t)jp;
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

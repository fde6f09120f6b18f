// dbaz_device.cuh -- device-side building blocks: bit-packed Dots & Boxes rules,
// NumPy-order reductions, plane writers.  sm_100a only.
//
// Reference behaviour restated here (paths relative to the reference root):
//   dots_boxes/dots_boxes_game.py:30-39   empty board, padding cells
//   dots_boxes/dots_boxes_game.py:44-49   get_valid_moves
//   dots_boxes/dots_boxes_game.py:51-59   get_result (early majority)
//   dots_boxes/dots_boxes_game.py:61-89   play_ (box closure, extra turn)
//   dots_boxes/dots_boxes_game.py:96-100  get_features
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "dbaz_b200.h"

namespace dbaz {

// Board geometry, passed to kernels by value (lives in the constant bank).
struct Board {
    int L, C, rows, cols, plane, A, nboxes, F;
    uint64_t real[2];  // bit a set iff action a is a real edge (not a padding cell)
};

// One child record of a tree node: 16 bytes so a lane moves it with one LDG.128 and the
// backup updates {W, N} with one 8-byte read-modify-write (mcts.py:55-62 keeps four
// separate per-child arrays; child_player_changed is recomputed from the edges instead).
struct __align__(16) Child {
    float W;        // child_total_value
    int32_t N;      // child_number_visits
    float prior;    // child_priors
    int32_t child;  // node index inside the tree's arena, 0 = not created yet
};

static_assert(sizeof(dbaz_state) == 32, "packed state must be 32 bytes");
static_assert(sizeof(Child) == 16, "child record must be 16 bytes");

enum : uint8_t { NF_EXPANDED = 1, NF_TERMINAL = 2 };

template <int NW>
struct Mask {
    uint64_t w[NW];
};

template <int NW>
__device__ __forceinline__ void mask_set(Mask<NW>& m, int a) {
    // (no m.w[a >> 6]: a run-time index puts the mask -- and whatever struct holds it -- into local memory)
    if (NW == 1) m.w[0] |= 1ull << a;
    else if (a < 64) m.w[0] |= 1ull << (a & 63);
    else m.w[NW - 1] |= 1ull << (a & 63);
}
template <int NW>
__device__ __forceinline__ bool mask_test(const Mask<NW>& m, int a) {
    if (NW == 1) return (m.w[0] >> a) & 1ull;
    return ((a < 64 ? m.w[0] : m.w[NW - 1]) >> (a & 63)) & 1ull;
}
template <int NW>
__device__ __forceinline__ bool mask_any(const Mask<NW>& m) {
    uint64_t r = 0;
#pragma unroll
    for (int i = 0; i < NW; ++i) r |= m.w[i];
    return r != 0;
}
// (e | own) contains every bit of box
template <int NW>
__device__ __forceinline__ bool mask_covers(const Mask<NW>& e, const Mask<NW>& box) {
    bool ok = true;
#pragma unroll
    for (int i = 0; i < NW; ++i) ok = ok && ((e.w[i] & box.w[i]) == box.w[i]);
    return ok;
}
template <int NW>
__device__ __forceinline__ Mask<NW> load_edges(const dbaz_state& s) {
    Mask<NW> m;
#pragma unroll
    for (int i = 0; i < NW; ++i) m.w[i] = s.edges[i];
    return m;
}
template <int NW>
__device__ __forceinline__ void store_edges(dbaz_state& s, const Mask<NW>& m) {
    s.edges[0] = m.w[0];
    s.edges[1] = NW == 2 ? m.w[NW - 1] : 0ull;
}

// The four edges of box (l, c): h(l,c), h(l+1,c), v(l,c), v(l,c+1)  (dots_boxes_game.py:102-104)
template <int NW>
__device__ __forceinline__ Mask<NW> box_mask(const Board& b, int l, int c) {
    Mask<NW> m;
#pragma unroll
    for (int i = 0; i < NW; ++i) m.w[i] = 0;
    mask_set(m, l * b.cols + c);
    mask_set(m, (l + 1) * b.cols + c);
    mask_set(m, b.plane + l * b.cols + c);
    mask_set(m, b.plane + l * b.cols + c + 1);
    return m;
}

// The (up to) two boxes an edge borders, in the order play_ tests them.  lc[j] = {l, c} or {-1,-1}.
template <int NW>
__device__ __forceinline__ void action_boxes(const Board& b, int a, Mask<NW> box[2], int lc[2][2]) {
    int p = a >= b.plane;
    int rem = a - p * b.plane;
    int l = rem / b.cols, c = rem - l * b.cols;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
#pragma unroll
        for (int i = 0; i < NW; ++i) box[j].w[i] = 0;
        lc[j][0] = lc[j][1] = -1;
    }
    if (p == 0) {
        if (c >= b.C) return;  // padding column
        if (l > 0) { box[0] = box_mask<NW>(b, l - 1, c); lc[0][0] = l - 1; lc[0][1] = c; }
        if (l < b.rows - 1) { box[1] = box_mask<NW>(b, l, c); lc[1][0] = l; lc[1][1] = c; }
    } else {
        if (l >= b.L) return;  // padding row
        if (c > 0) { box[0] = box_mask<NW>(b, l, c - 1); lc[0][0] = l; lc[0][1] = c - 1; }
        if (c < b.cols - 1) { box[1] = box_mask<NW>(b, l, c); lc[1][0] = l; lc[1][1] = c; }
    }
}

// Number of boxes that playing `a` closes, given the edges WITH a already set.
template <int NW>
__device__ __forceinline__ int closed_count(const Mask<NW>& e_after, const Mask<NW> box[2]) {
    int n = 0;
#pragma unroll
    for (int j = 0; j < 2; ++j) n += (mask_any(box[j]) && mask_covers(e_after, box[j])) ? 1 : 0;
    return n;
}

__device__ __forceinline__ int state_result(const dbaz_state& s) {
    if (s.btc2[0] == 0 && s.btc2[1] == 0) return 0;
    const int mine = s.to_play ? s.btc2[1] : s.btc2[0], other = s.to_play ? s.btc2[0] : s.btc2[1];
    if (mine < 0) return 1;
    if (other < 0) return -1;
    return DBAZ_RESULT_NONE;
}

// play_ on a packed state whose legality has been checked; n = boxes closed by the move.
template <int NW>
__device__ __forceinline__ void state_apply(dbaz_state& s, int a, int n_closed) {
    Mask<NW> e = load_edges<NW>(s);
    mask_set(e, a);
    store_edges<NW>(s, e);
    s.just_played = (int8_t)s.to_play;
    if (n_closed == 0) s.to_play = 1 - s.to_play;
    else if (s.to_play) s.btc2[1] = (int16_t)(s.btc2[1] - 2 * n_closed);
    else s.btc2[0] = (int16_t)(s.btc2[0] - 2 * n_closed);
}

template <int NW>
__device__ __forceinline__ bool state_legal(const Board& b, const dbaz_state& s, int a) {
    if (a < 0 || a >= b.A) return false;
    const bool lo = NW == 1 || a < 64;  // selects: a run-time index would put the state into local memory
    uint64_t real = lo ? b.real[0] : b.real[1], ed = lo ? s.edges[0] : s.edges[1];
    return ((real & ~ed) >> (a & 63)) & 1ull;
}

__device__ __forceinline__ void state_init(const Board& b, dbaz_state& s) {
    s.edges[0] = s.edges[1] = 0;
    s.btc2[0] = s.btc2[1] = (int16_t)b.nboxes;
    s.to_play = 0; s.just_played = -1; s.flags = 0; s.depth = 0; s.parent = -1; s.parent_action = -1;
    s.result = DBAZ_RESULT_NONE;
}

// ---- NumPy float add-reduce order for n <= 128 (pairwise_sum, unrolled by 8) ----
__device__ __forceinline__ float rn_add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double rn_add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float rn_div(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ double rn_div(double a, double b) { return __ddiv_rn(a, b); }

template <typename T>
__device__ __forceinline__ T np_sum(const T* a, int n) {
    if (n < 8) {
        T r = 0;
        for (int i = 0; i < n; ++i) r = rn_add(r, a[i]);
        return r;
    }
    T r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = a[j];
    int i = 8;
    for (; i < n - (n % 8); i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = rn_add(r[j], a[i + j]);
    }
    T res = rn_add(rn_add(rn_add(r[0], r[1]), rn_add(r[2], r[3])), rn_add(rn_add(r[4], r[5]), rn_add(r[6], r[7])));
    for (; i < n; ++i) res = rn_add(res, a[i]);
    return res;
}

// ---- board-plane writer (get_features + nn_batch_builder fused into the gather) ----
// Feature element e of a state: planes 0/1 = board//255, plane 2 = int8(2*boxes_to_close[to_play]).
template <int NW>
__device__ __forceinline__ float feature_value(const Board& b, const Mask<NW>& e, int k, int idx, int layout) {
    int ch, pos;
    if (layout == DBAZ_NCHW) { ch = idx / b.plane; pos = idx - ch * b.plane; }
    else { pos = idx / 3; ch = idx - pos * 3; }
    if (ch == 2) return (float)k;
    return mask_test(e, ch * b.plane + pos) ? 1.0f : 0.0f;
}

// Warp-cooperative: writes the F elements of one row.  When F % 4 == 0 every lane stores 4
// consecutive elements with one 16-byte (fp32) or 8-byte (16-bit types) vector store.
template <int NW>
__device__ __forceinline__ void write_planes_warp(const Board& b, const Mask<NW>& e, int k, void* planes, int64_t row,
                                                  int dtype, int layout, int lane) {
    const int F = b.F;
    if ((F & 3) == 0) {
        for (int q = lane; q < (F >> 2); q += 32) {
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = feature_value<NW>(b, e, k, 4 * q + j, layout);
            if (dtype == DBAZ_F32) {
                reinterpret_cast<float4*>(reinterpret_cast<float*>(planes) + row * F)[q] = make_float4(v[0], v[1], v[2], v[3]);
            } else {
                uint2 o;
                if (dtype == DBAZ_F16) {
                    __half2 a = __floats2half2_rn(v[0], v[1]), c = __floats2half2_rn(v[2], v[3]);
                    o.x = *reinterpret_cast<uint32_t*>(&a); o.y = *reinterpret_cast<uint32_t*>(&c);
                } else if (dtype == DBAZ_BF16) {
                    __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), c = __floats2bfloat162_rn(v[2], v[3]);
                    o.x = *reinterpret_cast<uint32_t*>(&a); o.y = *reinterpret_cast<uint32_t*>(&c);
                } else {
                    o.x = ((uint32_t)(uint16_t)(int16_t)v[0]) | ((uint32_t)(uint16_t)(int16_t)v[1] << 16);
                    o.y = ((uint32_t)(uint16_t)(int16_t)v[2]) | ((uint32_t)(uint16_t)(int16_t)v[3] << 16);
                }
                reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(planes) + row * F)[q] = o;
            }
        }
    } else {
        for (int i = lane; i < F; i += 32) {
            float v = feature_value<NW>(b, e, k, i, layout);
            if (dtype == DBAZ_F32) reinterpret_cast<float*>(planes)[row * F + i] = v;
            else if (dtype == DBAZ_F16) reinterpret_cast<__half*>(planes)[row * F + i] = __float2half_rn(v);
            else if (dtype == DBAZ_BF16) reinterpret_cast<__nv_bfloat16*>(planes)[row * F + i] = __float2bfloat16_rn(v);
            else reinterpret_cast<int16_t*>(planes)[row * F + i] = (int16_t)v;
        }
    }
}

// ---- Philox4x32-10 counter RNG for the rollout workload ----
__device__ __forceinline__ uint32_t philox_u32(uint64_t seed, uint64_t game, uint32_t ply) {
    uint32_t c0 = (uint32_t)game, c1 = (uint32_t)(game >> 32), c2 = ply, c3 = 0;
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
        c0 = n0; c1 = l1; c2 = n2; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return c0;
}

// index of the n-th (0-based) set bit of x; x must have more than n bits set
// (a search over population counts: six steps whatever n is, so the threads of a warp -- one game each, every one with its
// own n -- stay together; peeling the lowest set bit n times ran every warp for its largest n)
__device__ __forceinline__ int nth_set_bit(uint64_t x, int n) {
    uint32_t w = (uint32_t)x;
    int pos = 0;
    const int c0 = __popc(w);
    if (n >= c0) { n -= c0; w = (uint32_t)(x >> 32); pos = 32; }
    int off = 0;
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        const int c = __popc((w >> off) & ((1u << s) - 1u));
        if (n >= c) { n -= c; off += s; }
    }
    return pos + off;
}

}  // namespace dbaz

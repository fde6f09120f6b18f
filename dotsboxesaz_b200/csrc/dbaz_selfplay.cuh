// The per-check phase of the asynchronous self-play loop (BatchedSelfPlay.play_games_async): what the reference does
// between two searches of one game -- SelfPlay.get_next_move (self_play.py:27-49: temperature, draw of the move) and the
// body of SelfPlay.play_game (self_play.py:51-74: record the searched root, re-root, next search) -- for every tree whose
// search has finished, while the other trees keep searching.  Two kernels around the engine's own k_advance_roots:
//
//   k_selfplay_pick     one warp per tree, lane = action: visits -> (v / max)^(1/T) -> running sum -> the first action whose
//                       cumulative weight exceeds u * total (np.random.choice's searchsorted on the cdf, with the uniform
//                       pre-drawn per (move index, tree)); the searched root (state, visits, tree statistics, q) goes
//                       into the history row of its move index; moves[t] = the move, or -1 for trees that are not done.
//   k_selfplay_restart  after the re-root: games that go on get the simulation budget of their number of legal moves
//                       (n_searches) and their row of Dirichlet noise times the legal mask (mcts.py:220-223); the others
//                       are told to leave their tree alone (num_reads = -3, k_search_begin); counts the games still on.
#pragma once
#include "dbaz_tree_kernels.cuh"

namespace dbaz {

struct SelfplayArgs {
    int n_moves;
    const double* inv_temp;    // [n_moves]
    const double* uniforms;    // [n_moves][n]
    const double* noise;       // [n_moves][n][A] or null
    const int32_t* reads_by_k; // [A + 1]
    int8_t* searching;         // [n]
    int64_t* move_idx;         // [n]
    int32_t* moves;            // [n]
    dbaz_state* h_states;      // [n_moves][n]
    int32_t* h_visits;         // [n_moves][n][A]
    int8_t* h_active;          // [n_moves][n]
    int32_t* h_moves;          // [n_moves][n]
    int32_t* h_stats;          // [n_moves][n][8]
    float* h_q;                // [n_moves][n]
    double* noise_buf;         // [n][A]
    int32_t* reads;            // [n]
    int32_t* left;             // [1]
};

constexpr int SP_WARPS = 4;

__global__ void __launch_bounds__(SP_WARPS * 32) k_selfplay_pick(Board b, TreeArgs ta, SelfplayArgs sp) {
    const int lane = threadIdx.x & 31;
    const int t = blockIdx.x * SP_WARPS + (threadIdx.x >> 5);
    if (blockIdx.x == 0 && threadIdx.x == 0) *sp.left = 0;  // k_selfplay_restart (a later launch) counts into it
    if (t >= ta.n_trees) return;
    const TreeHot hot = load_hot(ta.trees + t);
    const bool done = sp.searching[t] && !(hot.sims_left > 0 || hot.n_pending > 0);
    if (!done) {
        if (lane == 0) sp.moves[t] = -1;
        return;
    }
    const int A = b.A, n = ta.n_trees;
    const int mi = (int)min((long long)sp.move_idx[t], (long long)sp.n_moves - 1);
    char* np = node_ptr(ta, t, 0);
    const dbaz_state h = load_hdr(np);
    const bool interior = (h.flags & NF_EXPANDED) && !(h.flags & NF_TERMINAL);
    const Child* ch = node_children(np);
    const int64_t row = (int64_t)mi * n + t;
    // ---- visits, their maximum
    int vis[DBAZ_MAX_ACTIONS / 32];
    int vmax = 0;
#pragma unroll
    for (int i = 0; i < DBAZ_MAX_ACTIONS / 32; ++i) {
        const int a = lane + 32 * i;
        vis[i] = (a < A && interior) ? ch[a].N : 0;
        vmax = max(vmax, vis[i]);
        if (a < A) sp.h_visits[row * A + a] = vis[i];
    }
    vmax = __reduce_max_sync(0xffffffffu, vmax);
    // ---- (v / max)^(1/T), running sum in action order, first action beyond u * total
    const double it = sp.inv_temp[mi];
    const double den = (double)max(vmax, 1);
    double carry = 0.0;
    double cdf[DBAZ_MAX_ACTIONS / 32];
    int last_pos = -1;  // the last action with a positive weight
#pragma unroll
    for (int i = 0; i < DBAZ_MAX_ACTIONS / 32; ++i) {
        const int a = lane + 32 * i;
        double p = vis[i] > 0 ? pow((double)vis[i] / den, it) : 0.0;
        if (a >= A) p = 0.0;
        const unsigned pos = __ballot_sync(0xffffffffu, p > 0.0);
        if (pos) last_pos = 32 * i + 31 - __clz(pos);
        double s = p;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const double o = __shfl_up_sync(0xffffffffu, s, d);
            if (lane >= d) s += o;
        }
        cdf[i] = carry + s;
        carry = __shfl_sync(0xffffffffu, cdf[i], 31);
    }
    const double target = sp.uniforms[row] * carry;
    int move = -1;
#pragma unroll
    for (int i = 0; i < DBAZ_MAX_ACTIONS / 32; ++i) {
        const int a = lane + 32 * i;
        const unsigned beyond = __ballot_sync(0xffffffffu, a < A && cdf[i] > target);
        if (move < 0 && beyond) move = 32 * i + __ffs(beyond) - 1;
    }
    if (move < 0) move = last_pos;  // u * total rounded up to the total
    if (lane == 0) {
        dbaz_state pub = h;
        pub.flags = 0; pub.depth = 0; pub.parent = -1; pub.parent_action = -1; pub.result = (int16_t)state_result(h);
        sp.h_states[row] = pub;
        const TreeRec T = ta.trees[t];
        int32_t* o = sp.h_stats + row * 8;
        o[0] = T.root_N; o[1] = T.max_deepness - T.deepness_correction; o[2] = T.tree_size; o[3] = T.terminal_count;
        o[4] = (h.flags & NF_EXPANDED) ? 1 : 0; o[5] = (h.flags & NF_TERMINAL) ? 1 : 0; o[6] = T.n_nodes;
        o[7] = (int32_t)(T.flags & (TF_ERR_POOL | TF_ERR_MOVE));
        sp.h_q[row] = __fdiv_rn(T.root_W, (float)(1 + T.root_N));  // mcts.py:35
        sp.h_moves[row] = move;
        sp.h_active[row] = 1;
        sp.moves[t] = move;
        if (move < 0) sp.searching[t] = 0;  // a finished search without a single visit (cannot happen for a legal budget): the
                                            // game stops here instead of being picked up again at every check
    }
}

// first != 0: the start of a batch of games -- every tree begins its first search, no move has been drawn
template <int NW>
__global__ void __launch_bounds__(SP_WARPS * 32) k_selfplay_restart(Board b, TreeArgs ta, SelfplayArgs sp, int first) {
    const int lane = threadIdx.x & 31;
    const int t = blockIdx.x * SP_WARPS + (threadIdx.x >> 5);
    if (t >= ta.n_trees) return;
    const bool done = first || sp.moves[t] >= 0;
    if (!done) {
        if (lane == 0) {
            sp.reads[t] = -3;
            if (sp.searching[t]) atomicAdd(sp.left, 1);
        }
        return;
    }
    const int A = b.A, n = ta.n_trees;
    const long long next = first ? sp.move_idx[t] : sp.move_idx[t] + 1;
    const dbaz_state h = load_hdr(node_ptr(ta, t, 0));  // the new root
    const bool go_on = state_result(h) == DBAZ_RESULT_NONE && next < sp.n_moves;
    int k = 0;
    for (int a0 = 0; a0 < A; a0 += 32) {
        const int a = a0 + lane;
        const bool legal = a < A && state_legal<NW>(b, h, a);
        k += __popc(__ballot_sync(0xffffffffu, legal));
        if (go_on && sp.noise && a < A) {
            const int64_t o = (int64_t)t * A + a;
            sp.noise_buf[o] = legal ? sp.noise[((int64_t)next * n) * A + o] : 0.0;  // mcts.py:223: noise * valid_actions
        }
    }
    if (lane == 0) {
        sp.move_idx[t] = next;
        sp.searching[t] = go_on ? 1 : 0;
        sp.reads[t] = go_on ? sp.reads_by_k[k] : -3;
        if (go_on) atomicAdd(sp.left, 1);
    }
}

}  // namespace dbaz

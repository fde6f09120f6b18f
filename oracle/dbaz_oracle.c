/*
 * dbaz_oracle.c -- CPU restatement of the damlobster/DotsBoxesAZ self-play hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this file.  The product path is the CUDA library in
 * dotsboxesaz_b200/csrc and fails loudly when that library is missing.
 *
 * Parity status: PINNED DIFFERENTIALLY.  The reference's own tests hold no usable
 * golden vectors for this path, so the pin is tests/golden/*.json,
 * produced by tests/golden/make_golden.py which imports the real reference from
 * /root/reference and records its outputs; tests/test_oracle_golden.py checks
 * every function below against those recordings.
 *
 * Deliberately written the way the reference is written (byte boards, one heap
 * object per node with per-child arrays) and NOT the way the CUDA engine is
 * written (bit masks, pooled 16-byte child records), so that agreement between
 * the two is meaningful.  Citations are to files under /root/reference.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_MAX_A 128
#define ORC_NONE 2 /* get_result() is None */

typedef struct {
    int L, C;       /* boxes: BOARD_DIM = (L, C), dots_boxes_game.py:23 */
    int rows, cols; /* L+1, C+1 */
    int plane;      /* rows*cols */
    int A;          /* NB_ACTIONS = 2*plane, dots_boxes_game.py:26 */
    int nboxes;     /* NB_BOXES */
} orc_game;

typedef struct {
    uint8_t board[ORC_MAX_A]; /* ravel of uint8[2][L+1][C+1]; 0 free, 1 padding, 255 played */
    int32_t to_play;
    int32_t just_played; /* -1 == None */
    int32_t btc2[2];     /* 2 * boxes_to_close (the reference keeps floats x.5) */
    uint64_t hash_lo, hash_hi; /* sum of 1<<move, dots_boxes_game.py:106-109 */
    int32_t hash_btc2;   /* 2 * hash[1]; 0 for the initial (0, 0) hash */
    int32_t pad_;
} orc_state;

int orc_sizeof_state(void) { return (int)sizeof(orc_state); }

/* ------------------------------------------------------------------ game */

/* dots_boxes_game.py:21-28 */
void orc_game_init(orc_game *g, int L, int C)
{
    g->L = L; g->C = C; g->rows = L + 1; g->cols = C + 1;
    g->plane = g->rows * g->cols; g->A = 2 * g->plane; g->nboxes = L * C;
}

/* dots_boxes_game.py:30-39 */
void orc_state_init(const orc_game *g, orc_state *s)
{
    memset(s, 0, sizeof(*s));
    for (int c = 0; c < g->cols; ++c) s->board[g->plane + g->L * g->cols + c] = 1; /* board[1, l, :] = 1 */
    for (int r = 0; r < g->rows; ++r) s->board[r * g->cols + g->C] = 1;             /* board[0, :, c] = 1 */
    s->to_play = 0; s->just_played = -1;
    s->btc2[0] = s->btc2[1] = g->nboxes; /* 2 * (NB_BOXES/2) */
}

/* dots_boxes_game.py:44-49 */
void orc_valid_moves(const orc_game *g, const orc_state *s, uint8_t *out)
{
    for (int a = 0; a < g->A; ++a) out[a] = (s->board[a] == 0);
}

/* dots_boxes_game.py:51-59; ORC_NONE stands for None */
int orc_result(const orc_game *g, const orc_state *s)
{
    (void)g;
    if (s->btc2[0] == 0 && s->btc2[1] == 0) return 0;
    if (s->btc2[s->to_play] < 0) return 1;
    if (s->btc2[1 - s->to_play] < 0) return -1;
    return ORC_NONE;
}

/* dots_boxes_game.py:102-104 */
static int check_box(const orc_game *g, const orc_state *s, int l, int c)
{
    int sum = s->board[l * g->cols + c] + s->board[(l + 1) * g->cols + c] +
              s->board[g->plane + l * g->cols + c] + s->board[g->plane + l * g->cols + c + 1];
    return sum == 4 * 255;
}

/* dots_boxes_game.py:61-89.  Returns number of closed boxes (0..2) or -1 for the
 * ValueError; closed_lc receives up to two (l, c) pairs in the reference's order. */
int orc_play(const orc_game *g, orc_state *s, int move, int32_t *closed_lc)
{
    if (move < 0 || move >= g->A) return -1;
    int p = move / g->plane, l = (move % g->plane) / g->cols, c = move % g->cols;
    if (s->board[move] != 0) return -1;
    s->board[move] = 255;
    int n = 0;
    int32_t tmp[4];
    if (!closed_lc) closed_lc = tmp;
    if (p == 0) {
        if (l > 0 && check_box(g, s, l - 1, c)) { closed_lc[2 * n] = l - 1; closed_lc[2 * n + 1] = c; ++n; }
        if (l < g->rows - 1 && check_box(g, s, l, c)) { closed_lc[2 * n] = l; closed_lc[2 * n + 1] = c; ++n; }
    } else {
        if (c > 0 && check_box(g, s, l, c - 1)) { closed_lc[2 * n] = l; closed_lc[2 * n + 1] = c - 1; ++n; }
        if (c < g->cols - 1 && check_box(g, s, l, c)) { closed_lc[2 * n] = l; closed_lc[2 * n + 1] = c; ++n; }
    }
    s->just_played = s->to_play;
    if (n == 0) s->to_play = 1 - s->to_play;
    else s->btc2[s->to_play] -= 2 * n;
    /* _update_hash, after the turn switch */
    if (move < 64) { uint64_t o = s->hash_lo; s->hash_lo += 1ull << move; if (s->hash_lo < o) s->hash_hi++; }
    else s->hash_hi += 1ull << (move - 64);
    s->hash_btc2 = s->btc2[s->to_play];
    return n;
}

/* dots_boxes_game.py:96-100: board//255 and a constant plane 2*boxes_to_close[to_play] */
void orc_features(const orc_game *g, const orc_state *s, int16_t *out)
{
    for (int a = 0; a < g->A; ++a) out[a] = s->board[a] / 255;
    int8_t k = (int8_t)s->btc2[s->to_play]; /* np.full_like(..., dtype=np.int8) */
    for (int i = 0; i < g->plane; ++i) out[g->A + i] = k;
}

/* --------------------------------------------------------- counter RNG */
/* Philox4x32-10 (Salmon et al. 2011), used only by the random-rollout workload
 * (key = seed, counter = (game, ply)). */
static void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1)
{
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1, n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

uint32_t orc_philox_u32(uint64_t seed, uint64_t game, uint32_t ply)
{
    uint32_t c[4] = { (uint32_t)game, (uint32_t)(game >> 32), ply, 0 };
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    return c[0];
}

/* Uniformly random legal playout to terminal.  moves_out (may be NULL) gets the
 * move list; returns the number of plies.  Move = the floor(u32*k / 2^32)-th legal
 * action in increasing action id. */
int orc_random_rollout(const orc_game *g, orc_state *s, uint64_t seed, uint64_t game, int32_t *moves_out)
{
    int ply = 0;
    while (orc_result(g, s) == ORC_NONE) {
        int k = 0;
        for (int a = 0; a < g->A; ++a) k += (s->board[a] == 0);
        if (k == 0) break;
        uint32_t u = orc_philox_u32(seed, game, (uint32_t)ply);
        int pick = (int)(((uint64_t)u * (uint64_t)k) >> 32), mv = -1;
        for (int a = 0; a < g->A; ++a) if (s->board[a] == 0 && pick-- == 0) { mv = a; break; }
        orc_play(g, s, mv, NULL);
        if (moves_out) moves_out[ply] = mv;
        ++ply;
    }
    return ply;
}

/* ------------------------------------------------------------- fake NN */
/* Deterministic stand-in for the policy/value net (SURVEY.md 8a KAT definition):
 *   h = hash[0] & 0xffffffff; raw_i = float32((h*2654435761 + i*40503) mod 1024) + 1
 *   p = raw / raw.sum() (fp32); v = float32(((h mod 2001) - 1000) / 1000).
 * `user` may point to an int selecting the variant (tests/golden/make_golden.py). */
void orc_fake_nn(const orc_game *g, const orc_state *s, float *p, float *v, void *user)
{
    int kind = user ? *(const int *)user : 0;
    uint64_t h = s->hash_lo & 0xffffffffull;
    if (kind == 0) {
        float sum = 0.0f;
        for (int i = 0; i < g->A; ++i) {
            p[i] = (float)((h * 2654435761ull + (uint64_t)i * 40503ull) % 1024ull) + 1.0f;
            sum += p[i]; /* integers <= 1024*128: exact in fp32 in any order */
        }
        for (int i = 0; i < g->A; ++i) p[i] = p[i] / sum;
        *v = (float)(((double)(h % 2001ull) - 1000.0) / 1000.0);
    } else { /* kind 1: uniform prior 1/A, five-level value: maximises exact score ties */
        float u = 1.0f / (float)g->A;
        for (int i = 0; i < g->A; ++i) p[i] = u;
        *v = (float)(((double)((h * 31ull) % 5ull) - 2.0) / 2.0);
    }
}

/* --------------------------------------------------------------- mcts */

typedef void (*orc_nn_fn)(const orc_game *, const orc_state *, float *p, float *v, void *user);

/* mcts.py:39-65 (UCTNode).  A node's own N/W/sign live in its parent's arrays at
 * [move]; the first node's live in the tree (TreeRoot, mcts.py:21-36). */
typedef struct orc_node {
    orc_state st;
    int move;
    struct orc_node *parent; /* NULL: parent is the TreeRoot */
    int is_expanded, is_terminal, deepness;
    struct orc_node **children; /* lazily created, mcts.py:53-54 */
    double *priors;             /* values; priors_f64 tells which arithmetic produced them */
    int priors_f64;
    float *W;
    int32_t *N;
    int32_t *sign; /* child_player_changed, initialised to +1 (mcts.py:61-62) */
} orc_node;

typedef struct {
    orc_game g;
    orc_node *first;
    /* TreeRoot fields */
    float root_W;   /* fp32 in effect: python float 0.0 absorbs into float32 arrays */
    int64_t root_N;
    int root_sign;
    int deepness_correction, max_deepness, terminal_count;
    int64_t tree_size;
    int64_t sims_done, path_nodes; /* instrumentation for bench.py (mean path length) */
} orc_tree;

static orc_node *node_new(const orc_game *g, const orc_state *st, int move, orc_node *parent, int parent_deepness)
{
    orc_node *n = (orc_node *)calloc(1, sizeof(orc_node));
    n->st = *st; n->move = move; n->parent = parent;
    n->is_terminal = orc_result(g, st) != ORC_NONE;
    n->children = (orc_node **)calloc(g->A, sizeof(orc_node *));
    n->priors = (double *)calloc(g->A, sizeof(double));
    n->W = (float *)calloc(g->A, sizeof(float));
    n->N = (int32_t *)calloc(g->A, sizeof(int32_t));
    n->sign = (int32_t *)malloc(g->A * sizeof(int32_t));
    for (int a = 0; a < g->A; ++a) n->sign[a] = 1;
    n->deepness = parent_deepness + 1;
    return n;
}

static void node_free(const orc_game *g, orc_node *n)
{
    if (!n) return;
    for (int a = 0; a < g->A; ++a) node_free(g, n->children[a]);
    free(n->children); free(n->priors); free(n->W); free(n->N); free(n->sign); free(n);
}

/* mcts.py:156-160 */
orc_tree *orc_tree_new(int L, int C, const orc_state *root_state)
{
    orc_tree *t = (orc_tree *)calloc(1, sizeof(orc_tree));
    orc_game_init(&t->g, L, C);
    t->first = node_new(&t->g, root_state, -1, NULL, 0);
    return t;
}

void orc_tree_free(orc_tree *t)
{
    if (!t) return;
    node_free(&t->g, t->first);
    free(t);
}

static int64_t node_N(const orc_tree *t, const orc_node *n) { return n->parent ? n->parent->N[n->move] : t->root_N; }

/* NumPy's float add-reduce order for n <= 128 (pairwise_sum in loops_utils.h):
 * 8 running accumulators, combined ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), then the tail. */
static float np_sum_f32(const float *a, int n)
{
    if (n < 8) { float r = 0.0f; for (int i = 0; i < n; ++i) r += a[i]; return r; }
    volatile float r[8];
    int i;
    for (i = 0; i < 8; ++i) r[i] = a[i];
    for (i = 8; i < n - (n % 8); i += 8) for (int j = 0; j < 8; ++j) r[j] = r[j] + a[i + j];
    float res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res += a[i];
    return res;
}
static double np_sum_f64(const double *a, int n)
{
    if (n < 8) { double r = 0.0; for (int i = 0; i < n; ++i) r += a[i]; return r; }
    volatile double r[8];
    int i;
    for (i = 0; i < 8; ++i) r[i] = a[i];
    for (i = 8; i < n - (n % 8); i += 8) for (int j = 0; j < 8; ++j) r[j] = r[j] + a[i + j];
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res += a[i];
    return res;
}

/* mcts.py:91-99, all float64 over float32/int32 storage; compiled with
 * -ffp-contract=off so no FMA is formed. */
void orc_ucb_scores(const orc_tree *t, const orc_node *n, double cpuct, double cpuct_base, double *out)
{
    int64_t N = node_N(t, n);
    double pb_c = log(((double)N + cpuct_base + 1.0) / cpuct_base) + cpuct;
    double sq = sqrt((double)N);
    for (int a = 0; a < t->g.A; ++a) {
        double pb = pb_c * (sq / (double)(n->N[a] + 1));
        double prior_score = pb * n->priors[a];
        double value_score = (double)n->W[a] / (double)(1 + n->N[a]);
        value_score *= (double)n->sign[a];
        out[a] = prior_score + value_score;
    }
}

/* mcts.py:101-103; np.argmax returns the first maximal index */
static int best_child(const orc_tree *t, const orc_node *n, double cpuct, double cpuct_base)
{
    double sc[ORC_MAX_A];
    orc_ucb_scores(t, n, cpuct, cpuct_base, sc);
    int best = 0; double bv = 0.0;
    for (int a = 0; a < t->g.A; ++a) {
        double inv = (n->st.board[a] == 0) ? 0.0 : 1.0;
        double v = -1e12 * inv + sc[a];
        if (a == 0 || v > bv) { bv = v; best = a; }
    }
    return best;
}

/* select_leaf (mcts.py:105-114): the first half of one _search(), up to the point where the
 * reference awaits the net.  Virtual loss is subtracted from every node left behind. */
static int select_path(orc_tree *t, double cpuct, double cpuct_base, orc_node **path)
{
    const orc_game *g = &t->g;
    int np_ = 0;
    orc_node *cur = t->first;
    path[np_++] = cur;
    while (cur->is_expanded && !cur->is_terminal) {
        /* current.total_value -= VIRTUAL_LOSS */
        if (cur->parent) cur->parent->W[cur->move] = cur->parent->W[cur->move] - 1.0f;
        else t->root_W = t->root_W - 1.0f;
        int best = best_child(t, cur, cpuct, cpuct_base);
        if (!cur->children[best]) {
            orc_state ns = cur->st;
            orc_play(g, &ns, best, NULL);
            cur->children[best] = node_new(g, &ns, best, cur, cur->deepness);
        }
        cur = cur->children[best];
        path[np_++] = cur;
    }
    return np_;
}

/* the second half of _search() (mcts.py:186-199): evaluate the leaf, mask/renormalise, expand, backup */
static void finish_path(orc_tree *t, orc_nn_fn nn, void *user, orc_node **path, int np_)
{
    const orc_game *g = &t->g;
    orc_node *leaf = path[np_ - 1];
    float value;
    if (!leaf->is_terminal) {
        float p[ORC_MAX_A];
        nn(g, &leaf->st, p, &value, user);
        for (int a = 0; a < g->A; ++a) p[a] = (leaf->st.board[a] == 0) ? p[a] : p[a] * 0.0f;
        float s = np_sum_f32(p, g->A);
        if (s > 0.0f && s != 1.0f) for (int a = 0; a < g->A; ++a) p[a] = p[a] / s;
        for (int a = 0; a < g->A; ++a) leaf->priors[a] = (double)p[a];
        leaf->priors_f64 = 0;
    } else {
        for (int a = 0; a < g->A; ++a) leaf->priors[a] = 0.0; /* np.zeros(A), float64 */
        leaf->priors_f64 = 1;
        value = (float)orc_result(g, &leaf->st);
    }
    /* expand */
    leaf->is_expanded = 1;
    int sgn = (leaf->st.to_play == leaf->st.just_played) ? 1 : -1;
    if (leaf->parent) leaf->parent->sign[leaf->move] = sgn; else t->root_sign = sgn;
    /* backup */
    for (int i = 0; i < np_; ++i) {
        orc_node *n = path[i];
        float v = (n->st.to_play == leaf->st.to_play) ? value : -value;
        float add = v + 1.0f; /* v + VIRTUAL_LOSS in fp32 */
        if (n->parent) { n->parent->W[n->move] = n->parent->W[n->move] + add; n->parent->N[n->move] += 1; }
        else { t->root_W = t->root_W + add; t->root_N += 1; }
    }
    t->terminal_count += leaf->is_terminal;
    if (leaf->deepness > t->max_deepness) t->max_deepness = leaf->deepness;
    t->sims_done += 1; t->path_nodes += np_;
}

/* One _search() (mcts.py:184-199), strictly sequential (max_pending_evals == 1). */
static void one_search(orc_tree *t, orc_nn_fn nn, void *user, double cpuct, double cpuct_base)
{
    orc_node *path[ORC_MAX_A + 2];
    int np_ = select_path(t, cpuct, cpuct_base, path);
    finish_path(t, nn, user, path, np_);
}

/* UCT_search (mcts.py:183-244); max_pending == 1 is the strictly sequential case.
 * noise == NULL <=> alpha <= 0 (then noise is the scalar 0.0 in the reference).
 * noise (float64[A]) is the host-drawn Dirichlet sample already multiplied by
 * the legal mask (mcts.py:220-223). */
int orc_uct_search_k(orc_tree *t, int num_reads, orc_nn_fn nn, void *user,
                     double cpuct, double cpuct_base, const double *noise, double coeff, int max_pending);

int orc_uct_search(orc_tree *t, int num_reads, orc_nn_fn nn, void *user,
                   double cpuct, double cpuct_base, const double *noise, double coeff)
{
    return orc_uct_search_k(t, num_reads, nn, user, cpuct, cpuct_base, noise, coeff, 1);
}

int orc_uct_search_k(orc_tree *t, int num_reads, orc_nn_fn nn, void *user,
                     double cpuct, double cpuct_base, const double *noise, double coeff, int max_pending)
{
    const int A = t->g.A;
    orc_node *root = t->first;
    if (!nn) nn = orc_fake_nn;
    if (!root->is_expanded) one_search(t, nn, user, cpuct, cpuct_base);

    /* mcts.py:213-226 */
    double probs[ORC_MAX_A];
    int probs_f64;
    if (root->priors_f64) {
        double s = np_sum_f64(root->priors, A);
        if (s != 0.0) { double s2 = np_sum_f64(root->priors, A); for (int a = 0; a < A; ++a) probs[a] = root->priors[a] / s2; }
        else for (int a = 0; a < A; ++a) probs[a] = 0.0;
        probs_f64 = 1;
    } else {
        float pf[ORC_MAX_A];
        for (int a = 0; a < A; ++a) pf[a] = (float)root->priors[a];
        float s = np_sum_f32(pf, A);
        if (s != 0.0f) { for (int a = 0; a < A; ++a) probs[a] = (double)(pf[a] / s); probs_f64 = 0; }
        else { for (int a = 0; a < A; ++a) probs[a] = 0.0; probs_f64 = 1; } /* np.zeros(A) is float64 */
    }
    if (probs_f64) {
        for (int a = 0; a < A; ++a) root->priors[a] = (1.0 - coeff) * probs[a] + coeff * (noise ? noise[a] : 0.0);
        root->priors_f64 = 1;
    } else if (noise) {
        float c1 = (float)(1.0 - coeff); /* python float is weak: float32 multiply */
        for (int a = 0; a < A; ++a) root->priors[a] = (double)(c1 * (float)probs[a]) + coeff * noise[a];
        root->priors_f64 = 1;
    } else {
        float c1 = (float)(1.0 - coeff), z = (float)(coeff * 0.0);
        for (int a = 0; a < A; ++a) root->priors[a] = (double)((c1 * (float)probs[a]) + z);
        root->priors_f64 = 0;
    }

    if (max_pending <= 1) {
        for (int i = 0; i < num_reads; ++i) one_search(t, nn, user, cpuct, cpuct_base);
        return 0;
    }
    /* mcts.py:228-242 with a net that suspends each _search() exactly once: the event loop then runs the
     * simulations in waves -- `max_pend` select_leaf()s back to back (each leaving its virtual loss behind; a terminal
     * leaf is expanded and backed up at once because it never awaits), then the evaluations / expands / backups of the
     * others in the same order -- the first wave min(max_pending_evals, A) wide, later ones max_pending_evals wide,
     * the last one whatever is left. */
    {
        static orc_node *paths[ORC_MAX_A][ORC_MAX_A + 2];
        int lens[ORC_MAX_A];
        int cap = max_pending < A ? max_pending : A, done = 0;
        if (cap > ORC_MAX_A) cap = ORC_MAX_A;
        while (done < num_reads) {
            int w = num_reads - done < cap ? num_reads - done : cap;
            for (int k = 0; k < w; ++k) {
                lens[k] = select_path(t, cpuct, cpuct_base, paths[k]);
                /* a terminal leaf never awaits the net (mcts.py:186,194-196): its _search() runs to completion on the spot */
                if (paths[k][lens[k] - 1]->is_terminal) { finish_path(t, nn, user, paths[k], lens[k]); lens[k] = 0; }
            }
            for (int k = 0; k < w; ++k) if (lens[k]) finish_path(t, nn, user, paths[k], lens[k]);
            done += w;
            cap = max_pending > ORC_MAX_A ? ORC_MAX_A : max_pending;
        }
    }
    return 0;
}

/* init_mcts_tree (mcts.py:163-180) */
int orc_reroot(orc_tree *t, int move, int reuse)
{
    const orc_game *g = &t->g;
    orc_node *prev = t->first;
    if (move < 0 || move >= g->A) return -1;
    if (!prev->children[move]) {
        orc_state ns = prev->st;
        if (orc_play(g, &ns, move, NULL) < 0) return -1;
        prev->children[move] = node_new(g, &ns, move, prev, prev->deepness);
    }
    orc_node *next = prev->children[move];
    int64_t nb_visits = prev->N[move];
    prev->children[move] = NULL;
    node_free(g, prev);
    t->root_W = 0.0f; t->root_N = 0; t->root_sign = 0;
    t->max_deepness = 0; t->terminal_count = 0;
    if (reuse) {
        next->parent = NULL;
        t->first = next;
        t->deepness_correction = next->deepness;
        t->tree_size = nb_visits;
    } else {
        orc_state st = next->st;
        node_free(g, next);
        t->first = node_new(g, &st, move, NULL, 0);
        t->deepness_correction = 0;
        t->tree_size = 0;
    }
    return 0;
}

/* ----------------------------------------------------------- accessors */
void orc_root_arrays(const orc_tree *t, int32_t *N, float *W, double *priors, int32_t *sign)
{
    const orc_node *r = t->first;
    for (int a = 0; a < t->g.A; ++a) {
        if (N) N[a] = r->N[a];
        if (W) W[a] = r->W[a];
        if (priors) priors[a] = r->priors[a];
        if (sign) sign[a] = r->sign[a];
    }
}
void orc_root_state(const orc_tree *t, orc_state *out) { *out = t->first->st; }
int orc_root_flags(const orc_tree *t) { return t->first->is_expanded | (t->first->is_terminal << 1) | (t->first->priors_f64 << 2); }
int64_t orc_root_N(const orc_tree *t) { return t->root_N; }
float orc_root_W(const orc_tree *t) { return t->root_W; }
/* TreeRoot.get_tree_stats (mcts.py:33-36): out = {max_deepness, tree_size, terminal_count}; returns q */
float orc_tree_stats(const orc_tree *t, int64_t *out)
{
    out[0] = t->max_deepness - t->deepness_correction;
    out[1] = t->tree_size;
    out[2] = t->terminal_count;
    return t->root_W / (float)(1 + t->root_N);
}
void orc_tree_counters(const orc_tree *t, int64_t *out) { out[0] = t->sims_done; out[1] = t->path_nodes; }
void orc_root_ucb(const orc_tree *t, double cpuct, double cpuct_base, double *out) { orc_ucb_scores(t, t->first, cpuct, cpuct_base, out); }

/* Bulk helper for parity runs and the CPU baseline: play `n_moves` argmax-visit
 * moves with `num_reads` sims each (fake NN, no noise), writing the visit vector
 * of every searched root to visits_out[n_moves_done][A].  Returns moves searched. */
int orc_selfplay_argmax(int L, int C, const orc_state *start, int num_reads, int max_moves,
                        double cpuct, double cpuct_base, int32_t *visits_out, int32_t *moves_out, int64_t *counters,
                        int kind)
{
    orc_tree *t = orc_tree_new(L, C, start);
    int A = t->g.A, m = 0;
    while (!t->first->is_terminal && m < max_moves) {
        orc_uct_search(t, num_reads, orc_fake_nn, &kind, cpuct, cpuct_base, NULL, 0.0);
        int best = 0;
        for (int a = 0; a < A; ++a) {
            visits_out[m * A + a] = t->first->N[a];
            if (t->first->N[a] > t->first->N[best]) best = a;
        }
        moves_out[m++] = best;
        if (counters) { counters[0] += t->sims_done; counters[1] += t->path_nodes; t->sims_done = t->path_nodes = 0; }
        orc_reroot(t, best, 1);
    }
    orc_tree_free(t);
    return m;
}

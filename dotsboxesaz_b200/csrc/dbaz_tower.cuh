// dbaz_tower.cuh -- host-side interface of the fused residual-tower kernel (dbaz_tower.cu), shared with dbaz_capi.cu.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>
#include <string>

namespace dbaz {

// Geometry of one CTA tile of the tower kernel for a board of H x W points (H = L + 1, W = C + 1).
// Activations of a tile live in shared memory as eight channel-group planes [cg][h][j][8 channels] (16 bytes per
// entry), j = board_in_tile * WP + w with WP = W + 1: column W of every board is a zero pad, so a shift of one row
// (one tap in x) never reads a neighbouring board; 128 rows per h-block = one tcgen05 M tile.
struct TowerGeom {
    int H, W, WP;
    int nb;          // boards per tile = 128 / WP
    int plane;       // bytes per channel-group plane: H * 128 * 16 + 128 (the last 128 bytes stay zero)
    int buf;         // bytes per activation buffer: 128 (zero) + 8 planes
    int ok;          // 1 if the kernel supports this board (2 <= H <= 6, WP <= 128)
};
TowerGeom tower_geom(int H, int W);

constexpr int TOWER_C = 64;                 // channels of the tower (nn.py / configuration.py: nb_channels = 64)
constexpr int TOWER_CHUNK_BYTES = 6144;     // one weight chunk: [2 k-planes][192 rows = 3 dy taps x 64 out channels][8 in channels] bf16
constexpr int TOWER_CHUNKS_PER_STAGE = 12;  // (3 dx) x (4 k-steps of 16 input channels)
constexpr int TOWER_HEAD_MAX = 64;          // output channels of the fused 1x1 head convolution (multiple of 16)

struct TowerLaunch {
    const void* blob;        // n_tiles * geom.buf bytes, planar tiles (tower_planarize / stem), pads zero
    const void* packed_w;    // n_stages * 12 chunks of 6144 bytes, then (if head_cout) one chunk of head_cout * 128 bytes
    const float* bias;       // [n_stages (+1)][64] float32 (BatchNorm folded)
    void* out;               // [n][H][W][cout_last] bf16, cout_last = head_cout ? head_cout : 64
    int n_stages;            // 3x3 conv stages (2 per residual block; stage s odd adds the block input and is in place)
    int head_cout;           // 0 or the channels of the fused 1x1 conv + ReLU that follows the tower
    int64_t n_boards;
    int n_sms;
    int* err_flag;           // device int: set non-zero if a barrier wait timed out
    long long* dbg;          // optional device int64[64][16]: clock64 timeline of CTA 0's first tile (diagnostics)
};

// Launches k_resnet_tower on `st`.  Returns an empty string on success, else the error text.
std::string tower_launch(const TowerGeom& g, const TowerLaunch& a, cudaStream_t st);
// [n][H][W][64] bf16 (NHWC) -> planar tiles
std::string tower_planarize(const TowerGeom& g, const void* nhwc, void* blob, int64_t n, cudaStream_t st);

}  // namespace dbaz

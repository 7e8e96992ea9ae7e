// hz_device.h -- interface between the host side of libhorizonator (hz_api.cpp, hz_dem.cpp) and
// its CUDA kernels (hz_kernels.cu).  Internal; the public C ABI is include/*.h.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

// ---- visibility key --------------------------------------------------------------------------
// One 64-bit word per pixel, resolved with atomicMin:
//     [63:40] q    window depth quantised to 24 bits (the GL_DEPTH_COMPONENT renderbuffer of
//                  horizonator-lib.c:646; smaller = nearer)
//     [39: 8] id   triangle number in the reference's draw order (horizonator-lib.c:496-508):
//                  2*(j*(2R-1)+i) + {0: (j,i),(j+1,i+1),(j+1,i)   1: (j,i),(j,i+1),(j+1,i+1)}
//     [ 7: 0] r8   the fragment's red channel (fragment.glsl:16 after unorm8 conversion)
// min() over keys == GL_LESS with in-order drawing: nearest q wins, equal q -> first drawn wins.
// The result is independent of the order in which threads get to a pixel.
#define HZ_KEY_CLEAR 0xFFFFFFFFFFFFFFFFull   /* q = 0xFFFFFF = cleared depth 1.0 */
#define HZ_Q_MAX     0xFFFFFFu

// ---- tiles as uploaded (raw file bytes) ------------------------------------------------------
struct HzTiles
{
    const uint8_t* tile[4][4];   // [i_lon][j_lat] device pointers to raw .hgt bytes, nullptr = 0
    int origin_cell[2];          // dem.h: origin_dem_cellij
    int ntiles[2];
    int cpd;                     // 1200 / 3600
};

// ---- one render ------------------------------------------------------------------------------
struct HzView
{
    // terrain (resident in HBM): N x N int16, row j = north index, column i = east index
    const int16_t* mosaic;
    int   N;                     // 2R
    int   pitch;                 // elements per row (multiple of 64)
    float* e_tab;                // [N] metres east of the eye for column i   (vertex.glsl:128-130)
    float* n_tab;                // [N] metres north of the eye for row j

    // eye
    float viewer_cell_i, viewer_cell_j, viewer_z;
    float deg_per_cell, cos_viewer_lat;

    // azimuth window; scalars the vertex shader derives from az_deg0/az_deg1 (vertex.glsl:139-150),
    // computed once on the host in float exactly as written there
    float az_center;             // az_rad_center
    float az_ndc_per_rad;        // 2/(az_rad1-az_rad0)
    float aspect;                // W/H of the FULL panorama

    float znear, zfar, znear_color, zfar_color;

    // target: the panorama is W x H; this render fills columns [x0, x1) of it
    int W, H;
    int x0, x1;
    unsigned long long* vis;     // [H][x1-x0], GL row order (row 0 = bottom)

    // triangles that passed k_march's integer tests (worst case: all of them)
    uint32_t* tri_queue;
    uint32_t* tri_count;
    // (triangle, band of rows) pairs too big for one thread of k_raster
    uint2*    big_queue;
    uint32_t* big_count;
    uint32_t  big_capacity;
    uint32_t* work_count;        // k_march's work-item dispenser

    // conservative culling of whole mesh blocks (never changes the image)
    float cull_d2_far;           // blocks entirely farther (horizontally) than sqrt(this) are skipped
    float cull_az_half;          // half-width [rad] of the az interval that can reach columns [x0,x1),
    float cull_az_mid;           //   centred here; cull_az_half >= pi disables the test
};

struct HzResolve
{
    const unsigned long long* vis;
    int   Wt, H;                 // target width (x1-x0), height
    const float* tanel;          // [H] tan(elevation) per GL row, host-computed (lib:1007-1012)
    float znear, zfar;
    uint8_t* image;              // [H][Wt][3] B,G,R top row first, or nullptr
    float*   ranges;             // [H][Wt] top row first, or nullptr
};

// All launches are asynchronous on `stream`.
cudaError_t hz_launch_mosaic (const HzTiles& t, int16_t* mosaic, int N, int pitch, cudaStream_t stream);
cudaError_t hz_launch_prepare(const HzView& v, cudaStream_t stream);   // clear vis, axis tables, queue
cudaError_t hz_launch_march  (const HzView& v, cudaStream_t stream);   // mesh + project + cull -> triangle list
cudaError_t hz_launch_raster (const HzView& v, cudaStream_t stream);   // set-up + rasterise the list
cudaError_t hz_launch_big    (const HzView& v, cudaStream_t stream);   // queued large triangles
cudaError_t hz_launch_resolve(const HzResolve& r, cudaStream_t stream);

// Column strip width of the march kernel, exposed for the tests/docs
constexpr int HZ_STRIP_CELLS = 62;
constexpr int HZ_SEG_ROWS    = 64;

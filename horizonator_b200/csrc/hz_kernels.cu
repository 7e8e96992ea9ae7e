// hz_kernels.cu -- sm_100a kernels of libhorizonator's render path.
//
//   k_mosaic   (init)    raw big-endian .hgt tiles -> one int16 mosaic        replaces dem.c:264-309 called
//                                                                              (2R)^2 times at horizonator-lib.c:435-439
//   k_prepare  (render)  clear visibility keys, per-column/row metre tables    glClear, lib:896; vertex.glsl:128-130
//   k_march    (render)  mesh generation + projection + exact integer cull       lib:496-508 (index pattern), vertex.glsl,
//                        -> list of triangles that can produce a fragment        GL cull/clip
//   k_raster   (render)  set-up + rasterisation + depth test of the survivors    vertex.glsl, geometry.glsl, GL raster,
//                        (one thread per triangle)                               depth test, fragment.glsl
//   k_big      (render)  the few triangles with large bounding boxes           same stages, one warp per band of rows
//   k_resolve  (render)  keys -> BGR8 image + float range image, top row first lib:936-1048
//
// The mesh is never materialised: triangle t of the reference's index buffer is (cell = t>>1, half = t&1)
// and its vertices are read straight from the int16 mosaic.
//
// Compiled with -fmad=false: plain a*b+c is two IEEE roundings (see hz_math.cuh).
#include "hz_device.h"
#include "hz_math.cuh"

#include <cstdint>

// ================================================================================================
// k_mosaic
// ================================================================================================

// dem.c:278-293 for one axis: mosaic index -> (tile, cell inside tile).  Cell 0 of a tile is read from the
// previous tile's last row/column (tiles overlap by one sample); with no previous tile the reference reads
// out of bounds, here it reads cell 0 of tile 0.
__device__ __forceinline__ void hz_split_cell(int g, int cpd, int& t, int& c)
{
    t = g / cpd;
    c = g - t * cpd;
    if(c == 0) { t--; c = cpd; }
    if(t < 0)  { t = 0; c = 0; }
}

__device__ __forceinline__ int16_t hz_decode_be16(const uint8_t* p)
{
    const int16_t z = (int16_t)(((unsigned)p[0] << 8) | (unsigned)p[1]);   // dem.c:307
    return z < 0 ? (int16_t)0 : z;                                          // dem.c:308
}

// Each thread produces 8 consecutive samples of one mosaic row (one 16-byte store).  When the 8 samples come
// from one tile row they are fetched with aligned 32-bit loads, byte-swapped with PRMT and clamped two at a
// time; otherwise (tile boundary) sample by sample.
__global__ void __launch_bounds__(128)
k_mosaic(const HzTiles T, int16_t* __restrict__ out, int N, int pitch)
{
    const int j  = blockIdx.y;
    const int i0 = (blockIdx.x * blockDim.x + threadIdx.x) * 8;
    if(i0 >= pitch) return;

    const int cpd = T.cpd;
    int tj, cj;
    hz_split_cell(j + T.origin_cell[1], cpd, tj, cj);
    const size_t row_off = (size_t)(cpd - cj) * (size_t)(cpd + 1);          // dem.c:300-304: north row first

    uint32_t w[4] = {0, 0, 0, 0};    // 8 little-endian int16, clamped

    int ti0, ci0;
    hz_split_cell(i0 + T.origin_cell[0], cpd, ti0, ci0);
    const bool one_run = (i0 + 8 <= N) && (ci0 >= 1) && (ci0 + 7 <= cpd) && (tj < T.ntiles[1]) && (ti0 < T.ntiles[0]);

    if(one_run)
    {
        const uint8_t* tile = T.tile[ti0][tj];
        if(tile != nullptr)
        {
            const uint8_t* src = tile + 2 * (row_off + (size_t)ci0);
            const uintptr_t a  = (uintptr_t)src;
            const uint32_t* s4 = (const uint32_t*)(a & ~(uintptr_t)3);
            uint32_t r[5];
            #pragma unroll
            for(int k = 0; k < 4; k++) r[k] = __ldg(s4 + k);
            if(a & 2)
            {
                r[4] = __ldg(s4 + 4);
                #pragma unroll
                for(int k = 0; k < 4; k++) r[k] = __funnelshift_r(r[k], r[k + 1], 16);
            }
            #pragma unroll
            for(int k = 0; k < 4; k++)
            {
                const uint32_t sw = __byte_perm(r[k], 0, 0x2301);           // swap bytes inside each half
                w[k] = __vmaxs2(sw, 0u);                                     // per-half max(z, 0)
            }
        }
    }
    else
    {
        #pragma unroll
        for(int k = 0; k < 8; k++)
        {
            const int i = i0 + k;
            int16_t z = 0;
            if(i < N)
            {
                int ti, ci;
                hz_split_cell(i + T.origin_cell[0], cpd, ti, ci);
                if(ti >= T.ntiles[0] || tj >= T.ntiles[1]) z = -1;          // dem.c:293
                else
                {
                    const uint8_t* tile = T.tile[ti][tj];
                    if(tile != nullptr) z = hz_decode_be16(tile + 2 * (row_off + (size_t)ci));
                }
            }
            w[k >> 1] |= (uint32_t)(uint16_t)z << (16 * (k & 1));
        }
    }
    *(uint4*)(out + (size_t)j * pitch + i0) = make_uint4(w[0], w[1], w[2], w[3]);
}

cudaError_t hz_launch_mosaic(const HzTiles& t, int16_t* mosaic, int N, int pitch, cudaStream_t stream)
{
    dim3 block(128), grid((pitch / 8 + 127) / 128, N);
    k_mosaic<<<grid, block, 0, stream>>>(t, mosaic, N, pitch);
    return cudaGetLastError();
}

// ================================================================================================
// k_prepare
// ================================================================================================

__global__ void __launch_bounds__(256)
k_prepare(const __grid_constant__ HzView P)
{
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t nth = (size_t)gridDim.x * blockDim.x;

    // glClear: depth 1.0 everywhere.  Keys are cleared two at a time.
    const size_t nkeys = (size_t)P.H * (size_t)(P.x1 - P.x0);
    ulonglong2* v2 = (ulonglong2*)P.vis;
    for(size_t k = tid; k < nkeys / 2; k += nth) v2[k] = make_ulonglong2(HZ_KEY_CLEAR, HZ_KEY_CLEAR);
    if(tid == 0 && (nkeys & 1)) P.vis[nkeys - 1] = HZ_KEY_CLEAR;

    // vertex.glsl:128-130, operator by operator:
    //   e = (i - viewer_cell_i) * DEG_PER_CELL * Rearth * pi/180. * cos_viewer_lat
    //   n = (j - viewer_cell_j) * DEG_PER_CELL * Rearth * pi/180.
    for(size_t k = tid; k < (size_t)P.N; k += nth)
    {
        const float f = (float)(int)k;
        P.e_tab[k] = (f - P.viewer_cell_i) * P.deg_per_cell * HZ_REARTH_F * HZ_PI_F / 180.f * P.cos_viewer_lat;
        P.n_tab[k] = (f - P.viewer_cell_j) * P.deg_per_cell * HZ_REARTH_F * HZ_PI_F / 180.f;
    }
    if(tid == 0) { *P.big_count = 0; *P.tri_count = 0; *P.work_count = 0; }
}

cudaError_t hz_launch_prepare(const HzView& v, cudaStream_t stream)
{
    const size_t nkeys = (size_t)v.H * (size_t)(v.x1 - v.x0);
    size_t blocks = (nkeys / 2 + 255) / 256;
    if(blocks < (size_t)(v.N + 255) / 256) blocks = (v.N + 255) / 256;
    if(blocks > 148 * 16) blocks = 148 * 16;
    if(blocks < 1) blocks = 1;
    k_prepare<<<(unsigned)blocks, 256, 0, stream>>>(v);
    return cudaGetLastError();
}

// ================================================================================================
// projection and triangle set-up shared by k_march, k_raster and k_big
// ================================================================================================

#define HZ_GUARD_PX   2097152.0f      /* 2^21: triangles reaching beyond are dropped (oracle rule F5) */
#define HZ_SNAP_LIMIT 536870912.0f    /* 2^29 = guard band in 1/256 pixel */

struct HzVtx
{
    float xn, yn;    // gl_Position.x, gl_Position.y
    float z;         // terrain height, metres
};

// vertex.glsl:132-153 for one vertex at (e, n) metres from the eye with terrain height z
__device__ __forceinline__ void hz_project(const HzView& P, float e, float n, float z, HzVtx& v)
{
    const float h  = z - P.viewer_z;
    const float d2 = e * e + n * n;
    float az = hz_atan2_az(e, n);
    // unwrap_near_rad(az, az_rad_center), vertex.glsl:34-38; the division by 2*pi is a multiplication by
    // the rounded reciprocal here (the quotient only has to pick the right turn count)
    const float t = (az - P.az_center) * 0.15915494309189535f;
    az = (t - rintf(t)) * 2.f * HZ_PI_F + P.az_center;
    v.xn = (az - P.az_center) * P.az_ndc_per_rad;
    v.yn = hz_atan_el(h, d2) * P.aspect * P.az_ndc_per_rad;
    v.z  = z;
}

// window coordinate -> 1/256 pixel fixed point (oracle rule F2), saturated at the guard band so that the
// integer math downstream cannot overflow (triangles touching the guard band are dropped anyway)
__device__ __forceinline__ int hz_snap(float a)
{
    float t = a * 256.0f;
    if(fabsf(t) < 8388608.0f) t = floorf(t + 0.5f);
    t = fminf(fmaxf(t, -HZ_SNAP_LIMIT), HZ_SNAP_LIMIT);
    return (int)t;
}

// triangle number -> its three vertices (row j, column i), lib:496-508
__device__ __forceinline__ void hz_tri_vertices(unsigned int id, int N, int vj[3], int vi[3])
{
    const unsigned int cell = id >> 1;
    const int j = (int)(cell / (unsigned int)(N - 1)), i = (int)(cell % (unsigned int)(N - 1));
    vj[0] = j; vi[0] = i;
    if((id & 1u) == 0) { vj[1] = j + 1; vi[1] = i + 1; vj[2] = j + 1; vi[2] = i;     }
    else               { vj[1] = j;     vi[1] = i + 1; vj[2] = j + 1; vi[2] = i + 1; }
}

struct HzTri
{
    int X0, Y0, X1, Y1, X2, Y2;      // snapped window positions, 1/256 pixel
    int px0, px1, py0, py1;          // clipped pixel bounding box (inclusive)
    float xw0, yw0, xw1, yw1, xw2, yw2;
    // attribute planes through the unsnapped float vertices, anchored at vertex 0 (oracle rule F6)
    float z0w, r0, dzdx, dzdy, drdx, drdy;
    unsigned int id;
};

// geometry.glsl:21-27, guard band, back-face cull, bounding box.  False if the triangle produces nothing.
__device__ __forceinline__ bool
hz_tri_bounds(const HzView& P, const HzVtx& a, const HzVtx& b, const HzVtx& c, HzTri& T)
{
    const float xmax = fmaxf(fmaxf(a.xn, b.xn), c.xn);
    const float xmin = fminf(fminf(a.xn, b.xn), c.xn);
    if(xmax - xmin > 0.5f) return false;                                     // geometry.glsl:21-27

    const float halfW = 0.5f * (float)P.W, halfH = 0.5f * (float)P.H;
    T.xw0 = a.xn * halfW + halfW; T.yw0 = a.yn * halfH + halfH;              // viewport transform (F1)
    T.xw1 = b.xn * halfW + halfW; T.yw1 = b.yn * halfH + halfH;
    T.xw2 = c.xn * halfW + halfW; T.yw2 = c.yn * halfH + halfH;
    if(!(fabsf(T.xw0) < HZ_GUARD_PX && fabsf(T.yw0) < HZ_GUARD_PX &&
         fabsf(T.xw1) < HZ_GUARD_PX && fabsf(T.yw1) < HZ_GUARD_PX &&
         fabsf(T.xw2) < HZ_GUARD_PX && fabsf(T.yw2) < HZ_GUARD_PX)) return false;   // F5

    T.X0 = hz_snap(T.xw0); T.Y0 = hz_snap(T.yw0);
    T.X1 = hz_snap(T.xw1); T.Y1 = hz_snap(T.yw1);
    T.X2 = hz_snap(T.xw2); T.Y2 = hz_snap(T.yw2);

    // GL_CULL_FACE, front = counter-clockwise, y up (lib:184; F3)
    const long long area = (long long)(T.X1 - T.X0) * (T.Y2 - T.Y0) - (long long)(T.X2 - T.X0) * (T.Y1 - T.Y0);
    if(area <= 0) return false;

    const int bx0 = min(min(T.X0, T.X1), T.X2), bx1 = max(max(T.X0, T.X1), T.X2);
    const int by0 = min(min(T.Y0, T.Y1), T.Y2), by1 = max(max(T.Y0, T.Y1), T.Y2);
    // pixel centres sit at 256*p + 128
    T.px0 = max((bx0 + 127) >> 8, P.x0);
    T.px1 = min((bx1 - 128) >> 8, P.x1 - 1);
    T.py0 = max((by0 + 127) >> 8, 0);
    T.py1 = min((by1 - 128) >> 8, P.H - 1);
    return T.px0 <= T.px1 && T.py0 <= T.py1;
}

// vertex.glsl:155,159-160 for one vertex: window depth and red channel
__device__ __forceinline__ void hz_depth_shade(const HzView& P, float e, float n, float z, float& zw, float& r)
{
    const float h   = z - P.viewer_z;
    const float d2  = e * e + n * n;
    const float dne = sqrtf(d2);                                             // length(en)
    const float len = sqrtf(d2 + h * h);                                     // length(enh)
    const float zn  = (len - P.znear) / (P.zfar - P.znear) * 2.f - 1.f;
    zw = zn * 0.5f + 0.5f;                                                   // glDepthRange(0,1)
    r  = fmaxf(fminf((dne - P.znear_color) / (P.zfar_color - P.znear_color), 1.0f), 0.0f);
}

__device__ __forceinline__ void
hz_tri_planes(const HzView& P, HzTri& T,
              float e0, float n0, float z0, float e1, float n1, float z1, float e2, float n2, float z2)
{
    float zw0, zw1, zw2, r0, r1, r2;
    hz_depth_shade(P, e0, n0, z0, zw0, r0);
    hz_depth_shade(P, e1, n1, z1, zw1, r1);
    hz_depth_shade(P, e2, n2, z2, zw2, r2);

    const float ax = T.xw1 - T.xw0, ay = T.yw1 - T.yw0;
    const float bx = T.xw2 - T.xw0, by = T.yw2 - T.yw0;
    const float det = ax * by - bx * ay;
    const float inv = 1.0f / det;
    const float az = zw1 - zw0, bz = zw2 - zw0;
    const float ar = r1 - r0,   br = r2 - r0;
    T.dzdx = (az * by - bz * ay) * inv;
    T.dzdy = (bz * ax - az * bx) * inv;
    T.drdx = (ar * by - br * ay) * inv;
    T.drdy = (br * ax - ar * bx) * inv;
    T.z0w = zw0; T.r0 = r0;
}

// complete set-up of triangle `id` from the mosaic; false if it produces nothing
__device__ __forceinline__ bool hz_tri_setup(const HzView& P, unsigned int id, HzTri& T, bool with_planes)
{
    int vj[3], vi[3];
    hz_tri_vertices(id, P.N, vj, vi);
    float e[3], n[3];
    HzVtx v[3];
    #pragma unroll
    for(int k = 0; k < 3; k++)
    {
        e[k] = __ldg(P.e_tab + vi[k]);
        n[k] = __ldg(P.n_tab + vj[k]);
        const float z = (float)__ldg(P.mosaic + (size_t)vj[k] * P.pitch + vi[k]);
        hz_project(P, e[k], n[k], z, v[k]);
    }
    if(!hz_tri_bounds(P, v[0], v[1], v[2], T)) return false;
    T.id = id;
    if(with_planes) hz_tri_planes(P, T, e[0], n[0], v[0].z, e[1], n[1], v[1].z, e[2], n[2], v[2].z);
    return true;
}

// depth test + colour write for one covered pixel centre
__device__ __forceinline__ void hz_fragment(const HzView& P, const HzTri& T, int px, int py)
{
    const float cx = (float)px + 0.5f, cy = (float)py + 0.5f;
    const float ddx = cx - T.xw0, ddy = cy - T.yw0;
    const float zw = T.z0w + (T.dzdx * ddx + T.dzdy * ddy);
    if(!(zw >= 0.0f && zw <= 1.0f)) return;                                  // F5: view-volume clip per fragment
    const unsigned int q = (unsigned int)((double)zw * 16777215.0 + 0.5);   // F7
    if(q >= HZ_Q_MAX) return;                                                // cannot pass GL_LESS against 1.0
    float r = T.r0 + (T.drdx * ddx + T.drdy * ddy);
    r = fmaxf(fminf(r, 1.0f), 0.0f);
    const unsigned int r8 = (unsigned int)(r * 255.0f + 0.5f);              // F8
    const unsigned long long key = ((unsigned long long)q << 40) | ((unsigned long long)T.id << 8) | r8;
    atomicMin(&P.vis[(size_t)py * (size_t)(P.x1 - P.x0) + (size_t)(px - P.x0)], key);
}

// Edge functions E_k(P) = dx_k*(Py - Y_k) - dy_k*(Px - X_k) on the snapped positions (F3); an edge owns its
// boundary iff it runs downwards, or is horizontal running leftwards (F4).  I = int when every term fits in
// 32 bits (triangle smaller than 128 pixels), long long otherwise.
template <typename I>
struct HzEdges
{
    I dx0, dy0, dx1, dy1, dx2, dy2;
    I b0, b1, b2;
    __device__ __forceinline__ explicit HzEdges(const HzTri& T)
    {
        dx0 = (I)T.X1 - T.X0; dy0 = (I)T.Y1 - T.Y0;
        dx1 = (I)T.X2 - T.X1; dy1 = (I)T.Y2 - T.Y1;
        dx2 = (I)T.X0 - T.X2; dy2 = (I)T.Y0 - T.Y2;
        b0 = (dy0 < 0 || (dy0 == 0 && dx0 < 0)) ? 0 : 1;
        b1 = (dy1 < 0 || (dy1 == 0 && dx1 < 0)) ? 0 : 1;
        b2 = (dy2 < 0 || (dy2 == 0 && dx2 < 0)) ? 0 : 1;
    }
    __device__ __forceinline__ bool inside(const HzTri& T, int px, int py) const
    {
        const I Px = (I)px * 256 + 128, Py = (I)py * 256 + 128;
        const I E0 = dx0 * (Py - T.Y0) - dy0 * (Px - T.X0);
        const I E1 = dx1 * (Py - T.Y1) - dy1 * (Px - T.X1);
        const I E2 = dx2 * (Py - T.Y2) - dy2 * (Px - T.X2);
        return E0 >= b0 && E1 >= b1 && E2 >= b2;
    }
};

__device__ __forceinline__ bool hz_tri_is_small(const HzTri& T)
{
    const int bx0 = min(min(T.X0, T.X1), T.X2), bx1 = max(max(T.X0, T.X1), T.X2);
    const int by0 = min(min(T.Y0, T.Y1), T.Y2), by1 = max(max(T.Y0, T.Y1), T.Y2);
    return (bx1 - bx0) < 32768 && (by1 - by0) < 32768;      // every difference < 2^15 => products < 2^30
}

// ================================================================================================
// k_march: mesh generation + projection + conservative-exact triangle filter
// ================================================================================================
//
// One warp walks a strip of the mosaic northwards.  Lane l owns vertex columns c0+2l and c0+2l+1 (one aligned
// 32-bit load per row) and the two cells to their right; the right-hand neighbour's column arrives by
// shuffle, the previous row stays in registers, so every vertex is projected once per strip (plus one shared
// column between strips and one shared row between segments).  Only the snapped window position of a vertex
// is kept.  A cell whose snapped bounding box holds no pixel centre of the target is finished after a few
// integer instructions -- about 90% of them.  For the others both triangles get the same integer tests the
// rasteriser will apply (pixel centre inside the bounding box, inside the target, counter-clockwise), and
// the numbers of the survivors are appended to a global list through a per-warp shared-memory stage (one
// global atomic per ~64 survivors).  k_raster then sets those triangles up and draws them with every lane busy.
//
// Work items (strip x 64-row segment) are handed out through an atomic counter, ordered outwards from the
// eye: the expensive items next to the eye start first and the cheap far field fills in behind them.

#define HZ_WARPS_PER_CTA 8
#define HZ_STAGE_SLOTS   192       /* < 64 pending + at most 4*32 new per row */
#define HZ_STAGE_FLUSH   64
#define HZ_SMALL_MAX_PIX 8         /* k_raster draws bounding boxes up to this many pixels itself */
#define HZ_BAND_ROWS     8         /* bigger ones are cut into bands of rows for k_big */

struct HzLaneVtx { int X, Y; };

__device__ __forceinline__ HzLaneVtx hz_lane_vertex(const HzView& P, float e, float n, float z, float halfW, float halfH)
{
    HzVtx v;
    hz_project(P, e, n, z, v);
    HzLaneVtx o;
    o.X = hz_snap(v.xn * halfW + halfW);
    o.Y = hz_snap(v.yn * halfH + halfH);
    return o;
}

__device__ __forceinline__ HzLaneVtx hz_shfl_down1(const HzLaneVtx& v)
{
    HzLaneVtx o;
    o.X = __shfl_down_sync(0xffffffffu, v.X, 1);
    o.Y = __shfl_down_sync(0xffffffffu, v.Y, 1);
    return o;
}

// does the snapped bounding box of the four corners hold a pixel centre of the target?
__device__ __forceinline__ bool
hz_cell_alive(const HzView& P, const HzLaneVtx& a, const HzLaneVtx& b, const HzLaneVtx& c, const HzLaneVtx& d)
{
    const int bx0 = min(min(a.X, b.X), min(c.X, d.X)), bx1 = max(max(a.X, b.X), max(c.X, d.X));
    const int by0 = min(min(a.Y, b.Y), min(c.Y, d.Y)), by1 = max(max(a.Y, b.Y), max(c.Y, d.Y));
    const int px0 = (bx0 + 127) >> 8, px1 = (bx1 - 128) >> 8;
    const int py0 = (by0 + 127) >> 8, py1 = (by1 - 128) >> 8;
    return px0 <= px1 && py0 <= py1 && px1 >= P.x0 && px0 < P.x1 && py1 >= 0 && py0 < P.H;
}

// the rasteriser's integer tests for one triangle (hz_tri_bounds without the float-only seam/guard tests,
// which k_raster applies; a triangle that fails here cannot produce a fragment there)
__device__ __forceinline__ bool
hz_tri_alive(const HzView& P, const HzLaneVtx& a, const HzLaneVtx& b, const HzLaneVtx& c)
{
    const int bx0 = min(min(a.X, b.X), c.X), bx1 = max(max(a.X, b.X), c.X);
    const int by0 = min(min(a.Y, b.Y), c.Y), by1 = max(max(a.Y, b.Y), c.Y);
    const int px0 = (bx0 + 127) >> 8, px1 = (bx1 - 128) >> 8;
    const int py0 = (by0 + 127) >> 8, py1 = (by1 - 128) >> 8;
    if(!(px0 <= px1 && py0 <= py1 && px1 >= P.x0 && px0 < P.x1 && py1 >= 0 && py0 < P.H)) return false;
    const long long area = (long long)(b.X - a.X) * (c.Y - a.Y) - (long long)(c.X - a.X) * (b.Y - a.Y);
    return area > 0;
}

// conservative test: can any triangle of the block [c_lo..c_hi] x [r_lo..r_hi] (vertex indices) reach the target?
__device__ __forceinline__ bool hz_block_dead(const HzView& P, int c_lo, int c_hi, int r_lo, int r_hi)
{
    const float e_lo = __ldg(P.e_tab + c_lo), e_hi = __ldg(P.e_tab + c_hi);
    const float n_lo = __ldg(P.n_tab + r_lo), n_hi = __ldg(P.n_tab + r_hi);
    // nearest point of the rectangle to the eye
    const float ne = (e_lo > 0.f) ? e_lo : ((e_hi < 0.f) ? e_hi : 0.f);
    const float nn = (n_lo > 0.f) ? n_lo : ((n_hi < 0.f) ? n_hi : 0.f);
    const float d2min = ne * ne + nn * nn;
    // farther than zfar horizontally => slant range > zfar => window depth > 1 for every fragment
    if(d2min > P.cull_d2_far) return true;

    if(P.cull_az_half < 3.2f && (ne != 0.f || nn != 0.f))
    {
        // the rectangle does not contain the eye: its azimuths form an interval bounded by corner azimuths.
        // Measure every corner relative to the first one (the fan is < pi wide, so no wrap inside it) and the
        // first one relative to the middle of the interval of interest.
        const float ce[4] = { e_lo, e_hi, e_lo, e_hi };
        const float cn[4] = { n_lo, n_lo, n_hi, n_hi };
        const float a0 = atan2f(ce[0], cn[0]);
        float lo = 0.f, hi = 0.f;
        #pragma unroll
        for(int k = 1; k < 4; k++)
        {
            float d = atan2f(ce[k], cn[k]) - a0;
            d -= 6.28318530717958648f * rintf(d * 0.15915494309189535f);
            lo = fminf(lo, d); hi = fmaxf(hi, d);
        }
        float m = a0 - P.cull_az_mid;
        m -= 6.28318530717958648f * rintf(m * 0.15915494309189535f);
        const float margin = 1e-3f;
        const float blo = m + lo - margin, bhi = m + hi + margin;
        bool alive = false;
        #pragma unroll
        for(int turn = -1; turn <= 1; turn++)
        {
            const float s = 6.28318530717958648f * (float)turn;
            alive = alive || (blo + s <= P.cull_az_half && bhi + s >= -P.cull_az_half);
        }
        if(!alive) return true;
    }
    return false;
}

// k-th element of the sequence c, c+1, c-1, c+2, c-2, ... restricted to [0, n)
__device__ __forceinline__ int hz_outward(int c, int k, int n)
{
    const int m = min(c, n - 1 - c);
    if(k <= 2 * m) return (k & 1) ? c + (k + 1) / 2 : c - k / 2;
    return (n - 1 - c > c) ? c + (k - m) : c - (k - m);
}

__device__ __forceinline__ void hz_stage_flush(const HzView& P, unsigned int* stage, int& count, int lane)
{
    __syncwarp();
    if(count > 0)
    {
        unsigned int base = 0;
        if(lane == 0) base = atomicAdd(P.tri_count, (unsigned int)count);
        base = __shfl_sync(0xffffffffu, base, 0);
        for(int k = lane; k < count; k += 32) P.tri_queue[base + k] = stage[k];
        count = 0;
    }
    __syncwarp();
}

__global__ void __launch_bounds__(HZ_WARPS_PER_CTA * 32, 4)
k_march(const __grid_constant__ HzView P)
{
    __shared__ unsigned int s_stage[HZ_WARPS_PER_CTA][HZ_STAGE_SLOTS];

    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int N = P.N;
    const int n_strips = (N - 1 + HZ_STRIP_CELLS - 1) / HZ_STRIP_CELLS;
    const int n_segs   = (N - 1 + HZ_SEG_ROWS - 1) / HZ_SEG_ROWS;
    const unsigned int n_items = (unsigned int)n_strips * (unsigned int)n_segs;
    const int strip_eye = min(max((int)P.viewer_cell_i / HZ_STRIP_CELLS, 0), n_strips - 1);
    const int seg_eye   = min(max((int)P.viewer_cell_j / HZ_SEG_ROWS, 0), n_segs - 1);

    unsigned int* stage = s_stage[wib];
    int count = 0;
    const float halfW = 0.5f * (float)P.W, halfH = 0.5f * (float)P.H;

    for(;;)
    {
        unsigned int item = 0;
        if(lane == 0) item = atomicAdd(P.work_count, 1u);
        item = __shfl_sync(0xffffffffu, item, 0);
        if(item >= n_items) break;
        const int seg   = hz_outward(seg_eye,   (int)(item / (unsigned int)n_strips), n_segs);
        const int strip = hz_outward(strip_eye, (int)(item % (unsigned int)n_strips), n_strips);

        const int c0 = strip * HZ_STRIP_CELLS;                   // first vertex column of the strip (even)
        const int r0 = seg * HZ_SEG_ROWS;                        // first vertex row
        const int r1 = min(r0 + HZ_SEG_ROWS, N - 1);             // last vertex row (inclusive)
        const int c_last = min(c0 + HZ_STRIP_CELLS, N - 1);      // last vertex column any cell of the strip touches
        if(hz_block_dead(P, c0, c_last, r0, r1)) continue;

        // this lane's two vertex columns (clamped for loads; cells beyond the mesh are masked below)
        const int colA = c0 + 2 * lane, colB = colA + 1;
        const float eA = __ldg(P.e_tab + min(colA, N - 1)), eB = __ldg(P.e_tab + min(colB, N - 1));
        // cell 0 spans columns colA..colB, cell 1 spans colB..colA+2 (the neighbour lane's first column)
        const bool cell0_ok = (lane < 31) && (colB <= N - 1);
        const bool cell1_ok = (lane < 31) && (colA + 2 <= N - 1);

        // mosaic rows are pitch-aligned and colA is even: one 32-bit load fetches both heights
        const int16_t* mrow = P.mosaic + (size_t)r0 * P.pitch + min(colA, P.pitch - 2);

        HzLaneVtx pA, pB, pC;   // previous row
        {
            const unsigned int zz = __ldg((const unsigned int*)mrow);
            const float n = __ldg(P.n_tab + r0);
            pA = hz_lane_vertex(P, eA, n, (float)(short)(zz & 0xFFFFu), halfW, halfH);
            pB = hz_lane_vertex(P, eB, n, (float)(short)(zz >> 16),     halfW, halfH);
            pC = hz_shfl_down1(pA);
        }

        unsigned int zz_next = (r0 + 1 <= r1) ? __ldg((const unsigned int*)(mrow + P.pitch)) : 0u;
        float n_next = (r0 + 1 <= r1) ? __ldg(P.n_tab + r0 + 1) : 0.f;
        for(int j = r0 + 1; j <= r1; j++)
        {
            const unsigned int zz = zz_next;
            const float n = n_next;
            if(j + 1 <= r1)
            {
                zz_next = __ldg((const unsigned int*)(mrow + (size_t)(j + 1 - r0) * P.pitch));
                n_next  = __ldg(P.n_tab + j + 1);
            }
            const HzLaneVtx cA = hz_lane_vertex(P, eA, n, (float)(short)(zz & 0xFFFFu), halfW, halfH);
            const HzLaneVtx cB = hz_lane_vertex(P, eB, n, (float)(short)(zz >> 16),     halfW, halfH);
            const HzLaneVtx cC = hz_shfl_down1(cA);

            // cell (j-1, colA): corners pA pB / cA cB ; cell (j-1, colB): corners pB pC / cB cC
            unsigned int m = 0;
            if(cell0_ok && hz_cell_alive(P, pA, pB, cA, cB))
            {
                if(hz_tri_alive(P, pA, cB, cA)) m |= 1u;     // (j-1,i), (j,i+1), (j,i)
                if(hz_tri_alive(P, pA, pB, cB)) m |= 2u;     // (j-1,i), (j-1,i+1), (j,i+1)
            }
            if(cell1_ok && hz_cell_alive(P, pB, pC, cB, cC))
            {
                if(hz_tri_alive(P, pB, cC, cB)) m |= 4u;
                if(hz_tri_alive(P, pB, pC, cC)) m |= 8u;
            }
            if(__any_sync(0xffffffffu, m != 0))
            {
                const unsigned int id0 = 2u * ((unsigned int)(j - 1) * (unsigned int)(N - 1) + (unsigned int)colA);
                const unsigned int lt = (1u << lane) - 1u;
                #pragma unroll
                for(int b = 0; b < 4; b++)
                {
                    const bool on = (m >> b) & 1u;
                    const unsigned int ballot = __ballot_sync(0xffffffffu, on);
                    if(on) stage[count + __popc(ballot & lt)] = id0 + (unsigned int)b;
                    count += __popc(ballot);
                }
                if(count >= HZ_STAGE_FLUSH) hz_stage_flush(P, stage, count, lane);
            }
            pA = cA; pB = cB; pC = cC;
        }
    }
    hz_stage_flush(P, stage, count, lane);
}

cudaError_t hz_launch_march(const HzView& v, cudaStream_t stream)
{
    // persistent: 4 CTAs of 8 warps per SM, the warps pull (strip, segment) items until none are left
    k_march<<<148 * 4, HZ_WARPS_PER_CTA * 32, 0, stream>>>(v);
    return cudaGetLastError();
}

// ================================================================================================
// k_raster: one thread per surviving triangle
// ================================================================================================

template <typename I>
__device__ __forceinline__ void hz_draw_box(const HzView& P, const HzTri& T)
{
    const HzEdges<I> E(T);
    for(int py = T.py0; py <= T.py1; py++)
        for(int px = T.px0; px <= T.px1; px++)
            if(E.inside(T, px, py)) hz_fragment(P, T, px, py);
}

__global__ void __launch_bounds__(256)
k_raster(const __grid_constant__ HzView P)
{
    const unsigned int count = *P.tri_count;
    const unsigned int nth = gridDim.x * blockDim.x;
    for(unsigned int t = blockIdx.x * blockDim.x + threadIdx.x; t < count; t += nth)
    {
        HzTri T;
        const unsigned int id = P.tri_queue[t];
        if(!hz_tri_setup(P, id, T, true)) continue;
        const int bw = T.px1 - T.px0 + 1, bh = T.py1 - T.py0 + 1;
        if((long long)bw * bh > HZ_SMALL_MAX_PIX)
        {
            const unsigned int bands = (unsigned int)((bh + HZ_BAND_ROWS - 1) / HZ_BAND_ROWS);
            const unsigned int slot = atomicAdd(P.big_count, bands);
            if(slot + bands <= P.big_capacity)
            {
                for(unsigned int b = 0; b < bands; b++) P.big_queue[slot + b] = make_uint2(id, b);
                continue;
            }
            // queue full: draw it here (slow but correct)
        }
        if(hz_tri_is_small(T)) hz_draw_box<int>(P, T);
        else                   hz_draw_box<long long>(P, T);
    }
}

cudaError_t hz_launch_raster(const HzView& v, cudaStream_t stream)
{
    k_raster<<<148 * 8, 256, 0, stream>>>(v);
    return cudaGetLastError();
}

// ================================================================================================
// k_big: one warp per (triangle, band of rows), lanes spread over the band's pixels
// ================================================================================================

__global__ void __launch_bounds__(256)
k_big(const __grid_constant__ HzView P)
{
    unsigned int count = *P.big_count;
    if(count > P.big_capacity) count = P.big_capacity;
    const unsigned int lane = threadIdx.x & 31;
    const unsigned int nwarps = gridDim.x * (blockDim.x >> 5);
    for(unsigned int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); t < count; t += nwarps)
    {
        const uint2 entry = P.big_queue[t];
        HzTri T;
        if(!hz_tri_setup(P, entry.x, T, true)) continue;      // cannot happen: k_raster accepted it
        const int y0 = T.py0 + (int)entry.y * HZ_BAND_ROWS;
        const int y1 = min(y0 + HZ_BAND_ROWS - 1, T.py1);
        const int bw = T.px1 - T.px0 + 1;
        const int npix = bw * (y1 - y0 + 1);
        if(hz_tri_is_small(T))
        {
            const HzEdges<int> E(T);
            for(int p = (int)lane; p < npix; p += 32)
            {
                const int py = y0 + p / bw, px = T.px0 + p % bw;
                if(E.inside(T, px, py)) hz_fragment(P, T, px, py);
            }
        }
        else
        {
            const HzEdges<long long> E(T);
            for(int p = (int)lane; p < npix; p += 32)
            {
                const int py = y0 + p / bw, px = T.px0 + p % bw;
                if(E.inside(T, px, py)) hz_fragment(P, T, px, py);
            }
        }
    }
}

cudaError_t hz_launch_big(const HzView& v, cudaStream_t stream)
{
    k_big<<<148 * 4, 256, 0, stream>>>(v);
    return cudaGetLastError();
}

// ================================================================================================
// k_resolve
// ================================================================================================

// lib:1013-1025 for one pixel: depth as glReadPixels(GL_DEPTH_COMPONENT, GL_FLOAT) returns it -> range
__device__ __forceinline__ float hz_range_of_key(unsigned long long key, float tanel, float znear, float zfar)
{
    const unsigned int q = (unsigned int)(key >> 40);
    const float depth = (float)((double)q * (1.0 / 16777215.0));             // F7 read-back
    if(depth == 1.0f) return -1.0f;                                          // lib:1016
    const float length_en = depth * (zfar - znear) + znear;                  // lib:1018
    const float z = tanel * length_en;
    // hypotf (lib:1024): glibc evaluates it in double and rounds once
    return (float)sqrt((double)length_en * (double)length_en + (double)z * (double)z);
}

// 4 pixels per thread: 2x16 B of keys in, 12 B of BGR and 16 B of range out
__global__ void __launch_bounds__(256)
k_resolve4(const HzResolve R)
{
    const int groups_per_row = R.Wt >> 2;
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if(g >= (long long)groups_per_row * R.H) return;
    const int y  = (int)(g / groups_per_row);          // GL row (0 = bottom)
    const int x  = (int)(g % groups_per_row) << 2;
    const size_t src = (size_t)y * R.Wt + x;
    const size_t dst = (size_t)(R.H - 1 - y) * R.Wt + x;   // top row first (lib:949-958, 1026-1038)

    const ulonglong2 k01 = *(const ulonglong2*)(R.vis + src);
    const ulonglong2 k23 = *(const ulonglong2*)(R.vis + src + 2);
    const unsigned long long k[4] = { k01.x, k01.y, k23.x, k23.y };

    if(R.image)
    {
        // hit: (B,G,R) = (0,0,r8) ; sky: clear colour (0,0,1) read as BGR = (255,0,0)   lib:185, 938-939
        unsigned char b[12];
        #pragma unroll
        for(int p = 0; p < 4; p++)
        {
            const bool hit = (unsigned int)(k[p] >> 40) != HZ_Q_MAX;
            b[3 * p + 0] = hit ? 0 : 255;
            b[3 * p + 1] = 0;
            b[3 * p + 2] = hit ? (unsigned char)(k[p] & 0xFFu) : 0;
        }
        uint32_t* o = (uint32_t*)(R.image + dst * 3);      // dst*3 is a multiple of 4 because x and Wt are
        o[0] = b[0] | (b[1] << 8) | (b[2]  << 16) | ((uint32_t)b[3]  << 24);
        o[1] = b[4] | (b[5] << 8) | (b[6]  << 16) | ((uint32_t)b[7]  << 24);
        o[2] = b[8] | (b[9] << 8) | (b[10] << 16) | ((uint32_t)b[11] << 24);
    }
    if(R.ranges)
    {
        const float t = R.tanel[y];
        float4 r;
        r.x = hz_range_of_key(k[0], t, R.znear, R.zfar);
        r.y = hz_range_of_key(k[1], t, R.znear, R.zfar);
        r.z = hz_range_of_key(k[2], t, R.znear, R.zfar);
        r.w = hz_range_of_key(k[3], t, R.znear, R.zfar);
        *(float4*)(R.ranges + dst) = r;
    }
}

// any width
__global__ void __launch_bounds__(256)
k_resolve1(const HzResolve R)
{
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if(g >= (long long)R.Wt * R.H) return;
    const int y = (int)(g / R.Wt), x = (int)(g % R.Wt);
    const unsigned long long key = R.vis[g];
    const size_t dst = (size_t)(R.H - 1 - y) * R.Wt + x;
    if(R.image)
    {
        const bool hit = (unsigned int)(key >> 40) != HZ_Q_MAX;
        R.image[dst * 3 + 0] = hit ? 0 : 255;
        R.image[dst * 3 + 1] = 0;
        R.image[dst * 3 + 2] = hit ? (unsigned char)(key & 0xFFu) : 0;
    }
    if(R.ranges) R.ranges[dst] = hz_range_of_key(key, R.tanel[y], R.znear, R.zfar);
}

cudaError_t hz_launch_resolve(const HzResolve& r, cudaStream_t stream)
{
    const bool aligned = (r.Wt % 4 == 0) && (((uintptr_t)r.image & 3) == 0) && (((uintptr_t)r.ranges & 15) == 0);
    if(aligned)
    {
        const long long n = (long long)(r.Wt / 4) * r.H;
        k_resolve4<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(r);
    }
    else
    {
        const long long n = (long long)r.Wt * r.H;
        k_resolve1<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(r);
    }
    return cudaGetLastError();
}

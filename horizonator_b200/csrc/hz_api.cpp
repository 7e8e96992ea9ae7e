// hz_api.cpp -- the C ABI of include/horizonator.h and include/horizonator-batch.h.
//
// Replaces the host side of /root/reference/horizonator-lib.c: where that file drives an OpenGL context, this
// one owns a block of device state per context (the decoded DEM square, the visibility buffer, output and
// staging buffers, one CUDA stream) and enqueues the kernels of hz_kernels.cu.  The caller-visible struct has
// no room for a pointer, so the state lives in a process-wide table and `ctx->program` holds its handle.
//
// No CPU rendering path exists: without a usable CUDA device horizonator_init() fails.
#include "horizonator.h"
#include "horizonator-batch.h"
#include "util.h"
#include "hz_device.h"

#include <cmath>
#include <cstddef>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

namespace {

#define CUDA_TRY(expr)                                                              \
    do {                                                                            \
        cudaError_t e__ = (expr);                                                   \
        if(e__ != cudaSuccess)                                                      \
        {                                                                           \
            MSG("CUDA error: %s failed: %s", #expr, cudaGetErrorString(e__));       \
            return false;                                                           \
        }                                                                           \
    } while(0)

constexpr int   TANEL_SLOTS     = 8;
constexpr float PI_F            = 3.14159265358979f;       // vertex.glsl:31
// Queue capacities grow with the image (set in alloc_target); whatever does not fit is drawn by a slow in-kernel
// path, so these only have to be generous, not safe.
//   triangles of one stage awaiting set-up:                    max(2^22, pixels/2)
//   (record, sub-box) pairs for the large-triangle kernel:     max(2^21, pixels/4) per pass
//   set-up records of those triangles:                         max(2^18, pixels/16); beyond that k_big repeats the set-up
constexpr int   PROF_EVENTS     = 7;            // 6 stages per render
constexpr int   MAX_BANDS       = HZ_MAX_BANDS;
// [0] big_count near, [1] spare, [2] tri_count near, [3] big-triangle records, [4+4b] tile_count,
// [5+4b] block_count, [6+4b] tri_count, [7+4b] big_count of band b, then the stats
constexpr int   STATS_AT        = 4 + 4 * MAX_BANDS;
constexpr int   N_COUNTERS      = STATS_AT + HZ_STAT_COUNT;

// what the reference keeps in GL uniforms
struct ViewState
{
    float viewer_cell_i = 0, viewer_cell_j = 0, viewer_z = 0, cos_viewer_lat = 1;
    float az_deg0 = -45.f, az_deg1 = 45.f;
};

constexpr int PARAM_RING = 16;

// everything one render writes besides its outputs
struct Scratch
{
    cudaStream_t stream = nullptr;     // lanes only
    cudaEvent_t  done = nullptr;       // lanes only
    unsigned long long* d_vis = nullptr;
    float *d_e = nullptr, *d_n = nullptr;
    uint32_t *d_tile_queue = nullptr, *d_block_queue = nullptr, *d_tri_queue = nullptr;
    uint2*    d_big_queue = nullptr;   // [2][BIG_CAPACITY]: near pass, bands
    uint4*    d_bigtri = nullptr;      // [BIGTRI_CAPACITY][6]
    uint32_t* d_counters  = nullptr;   // [N_COUNTERS]
    uint8_t*  d_image  = nullptr;      // lanes only: staging of one view's outputs for the host-pointer batch call
    float*    d_ranges = nullptr;

    // parameters of the render in flight: the kernels read d_views; the host fills a slot of the pinned ring and
    // copies it over (the ring lets several renders be queued without waiting)
    HzView* d_views = nullptr;         // [HZ_V_COUNT]
    HzView* h_views = nullptr;         // pinned [PARAM_RING][HZ_V_COUNT]
    cudaEvent_t ring_ev[PARAM_RING] = {};
    int ring_next = 0;
    cudaEvent_t  busy = nullptr;       // recorded after the last render that used this set
    cudaStream_t last_stream = nullptr;
    bool busy_recorded = false;
    cudaGraphExec_t graph = nullptr;   // the standard chain, captured on first use
    int  graph_launches = 0;
    bool graph_big_per_band = false;   // the shape the captured chain has (see big_after_every_band)
    bool graph_failed = false;
};

struct Slot
{
    int device = 0;
    cudaStream_t stream = nullptr;

    // terrain
    int N = 0, pitch = 0, cpd = 0;
    int16_t* d_mosaic = nullptr;
    HzTiles  tiles{};
    short2 *d_mm_block = nullptr, *d_mm_tile = nullptr;   // culling pyramid
    int nb = 0, nt = 0;
    int near_rings = 2;
    int occl_tile_max_pix = 64, occl_block_max_pix = 32, small_max_pix = 16;
    int grid_percent_single = 150, grid_percent_batch = 35;   // see hz_grid() in hz_kernels.cu
    // Rings (in tiles around the eye's tile) at which the bands end; the last band runs to the edge of the mesh.
    // More bands = more of the mesh culled by what nearer bands drew, but four more kernels each.  A lone view is
    // latency-bound and gets two bands; the views of a batch overlap each other's latencies and get three (measured
    // over a grid of viewpoints: +18 % throughput, see profiles/).  The image is the same either way.
    struct Bands { int n; int end[MAX_BANDS]; };
    Bands bands_single = { 2, { 48, 1 << 20, 0, 0, 0, 0 } };
    Bands bands_batch  = { 3, { 24, 72, 1 << 20, 0, 0, 0 } };
    int n_lanes_max = 16;

    // target
    int W = 0, H = 0;
    uint8_t* d_image  = nullptr;
    float*   d_ranges = nullptr;
    size_t   target_pixels = 0;      // capacity of the buffers above
    uint32_t tri_capacity = 0, big_capacity = 0, bigtri_capacity = 0;

    // tan(elevation) per row, computed on the host like the reference's read-back does
    float*   d_tanel = nullptr;      // [TANEL_SLOTS][H]
    float*   h_tanel = nullptr;      // pinned, same shape
    struct { bool valid; float daz; int W, H; } tanel_key[TANEL_SLOTS] = {};
    int      tanel_next = 0;

    // Scratch of one render in flight.  `main` serves the single-view entry points (and keeps the last
    // visibility buffer for horizonator_pick); `lanes` are made on demand by the batch entry points so that
    // several views of one batch render concurrently, each on its own stream.
    Scratch main;
    std::vector<Scratch> lanes;
    cudaEvent_t fork_ev = nullptr;

    ViewState view;
    float znear = HORIZONATOR_ZNEAR_DEFAULT, zfar = HORIZONATOR_ZFAR_DEFAULT;
    float znear_color = HORIZONATOR_ZNEAR_DEFAULT, zfar_color = HORIZONATOR_ZFAR_DEFAULT;

    bool have_render = false;        // d_vis holds a complete full-width render (for pick)
    unsigned launches_last = 0;

    // optional per-kernel timing: PROF_EVENTS events per recorded render
    bool profiling = false;
    bool  seam_wrap = false;         // opt-in (horizonator_set_seam_wrap); false = seam triangles dropped like the reference
    float curvature = 0.f;           // opt-in (horizonator_set_earth_curvature); 0 = flat earth like the reference
    bool use_graphs = true;
    bool collect_stats = false;      // culling counters (horizonator_render_counters); off: the kernels skip them
    std::vector<cudaEvent_t> prof_events;
    size_t prof_used = 0;
};

std::mutex          g_table_mutex;
std::vector<Slot*>  g_table;         // handle = index + 1

Slot* slot_of(const horizonator_context_t* ctx)
{
    if(ctx == nullptr || ctx->Ntriangles <= 0 || ctx->program == 0) return nullptr;
    std::lock_guard<std::mutex> lock(g_table_mutex);
    if(ctx->program > g_table.size()) return nullptr;
    return g_table[ctx->program - 1];
}

struct DeviceGuard
{
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int dev)
    {
        if(cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        ok = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard() { if(prev >= 0) cudaSetDevice(prev); }
};

void drop_graph(Scratch& c)
{
    if(c.graph) cudaGraphExecDestroy(c.graph);
    c.graph = nullptr; c.graph_failed = false;
}

// the part of a scratch set whose size depends on the image
void free_scratch_target(Scratch& c)
{
    drop_graph(c);
    cudaFree(c.d_vis); cudaFree(c.d_image); cudaFree(c.d_ranges);
    cudaFree(c.d_tri_queue); cudaFree(c.d_big_queue); cudaFree(c.d_bigtri);
    c.d_vis = nullptr; c.d_image = nullptr; c.d_ranges = nullptr;
    c.d_tri_queue = nullptr; c.d_big_queue = nullptr; c.d_bigtri = nullptr;
}

bool alloc_scratch_target(const Slot& s, Scratch& c, bool staging)
{
    const size_t px = s.target_pixels;
    CUDA_TRY(cudaMalloc(&c.d_vis, px * sizeof(unsigned long long)));
    CUDA_TRY(cudaMalloc(&c.d_tri_queue, (size_t)s.tri_capacity * sizeof(uint32_t)));
    CUDA_TRY(cudaMalloc(&c.d_big_queue, 2 * (size_t)s.big_capacity * sizeof(uint2)));
    CUDA_TRY(cudaMalloc(&c.d_bigtri, (size_t)s.bigtri_capacity * 6 * sizeof(uint4)));
    if(staging)
    {
        CUDA_TRY(cudaMalloc(&c.d_image, px * 3));
        CUDA_TRY(cudaMalloc(&c.d_ranges, px * sizeof(float)));
    }
    return true;
}

void free_target(Slot& s)
{
    free_scratch_target(s.main);
    for(Scratch& l : s.lanes) free_scratch_target(l);
    cudaFree(s.d_image);  s.d_image = nullptr;
    cudaFree(s.d_ranges); s.d_ranges = nullptr;
    cudaFree(s.d_tanel);  s.d_tanel = nullptr;
    cudaFreeHost(s.h_tanel);  s.h_tanel = nullptr;
    s.target_pixels = 0;
    for(auto& k : s.tanel_key) k.valid = false;
}

bool alloc_target(Slot& s, int W, int H)
{
    free_target(s);
    const size_t px = (size_t)W * (size_t)H;
    auto cap = [px](size_t floor_, size_t div) {
        const size_t c = px / div > floor_ ? px / div : floor_;
        return (uint32_t)(c > 0x7FFFFFFFu ? 0x7FFFFFFFu : c);
    };
    s.target_pixels = px;
    s.tri_capacity = cap((size_t)1 << 22, 2); s.big_capacity = cap((size_t)1 << 21, 4); s.bigtri_capacity = cap((size_t)1 << 18, 16);
    // tests shrink the queues to exercise the overflow paths
    if(const char* env = getenv("HORIZONATOR_TRI_CAPACITY"))    s.tri_capacity    = (uint32_t)(atoi(env) > 1 ? atoi(env) : 1);
    if(const char* env = getenv("HORIZONATOR_BIG_CAPACITY"))    s.big_capacity    = (uint32_t)(atoi(env) > 1 ? atoi(env) : 1);
    if(const char* env = getenv("HORIZONATOR_BIGTRI_CAPACITY")) s.bigtri_capacity = (uint32_t)(atoi(env) > 1 ? atoi(env) : 1);
    if(!alloc_scratch_target(s, s.main, false)) return false;
    for(Scratch& l : s.lanes) if(!alloc_scratch_target(s, l, true)) return false;
    CUDA_TRY(cudaMalloc(&s.d_image, px * 3));
    CUDA_TRY(cudaMalloc(&s.d_ranges, px * sizeof(float)));
    CUDA_TRY(cudaMalloc(&s.d_tanel, (size_t)TANEL_SLOTS * H * sizeof(float)));
    CUDA_TRY(cudaMallocHost(&s.h_tanel, (size_t)TANEL_SLOTS * H * sizeof(float)));
    s.W = W; s.H = H; s.target_pixels = px;
    s.have_render = false;
    return true;
}

void free_scratch(Scratch& c)
{
    free_scratch_target(c);
    cudaFree(c.d_e); cudaFree(c.d_n);
    cudaFree(c.d_tile_queue); cudaFree(c.d_block_queue);
    cudaFree(c.d_counters);
    cudaFree(c.d_views); cudaFreeHost(c.h_views);
    for(cudaEvent_t e : c.ring_ev) if(e) cudaEventDestroy(e);
    if(c.busy) cudaEventDestroy(c.busy);
    if(c.done) cudaEventDestroy(c.done);
    if(c.stream) cudaStreamDestroy(c.stream);
    c = Scratch{};
}

// everything but the visibility buffer (alloc_target owns that: it depends on the image size)
bool alloc_scratch(const Slot& s, Scratch& c, bool own_stream)
{
    CUDA_TRY(cudaMalloc(&c.d_e, (size_t)s.N * sizeof(float)));
    CUDA_TRY(cudaMalloc(&c.d_n, (size_t)s.N * sizeof(float)));
    CUDA_TRY(cudaMalloc(&c.d_tile_queue, (size_t)s.nt * s.nt * sizeof(uint32_t)));
    CUDA_TRY(cudaMalloc(&c.d_block_queue, (size_t)s.nb * s.nb * sizeof(uint32_t)));
    CUDA_TRY(cudaMalloc(&c.d_counters, N_COUNTERS * sizeof(uint32_t)));
    CUDA_TRY(cudaMalloc(&c.d_views, HZ_V_COUNT * sizeof(HzView)));
    CUDA_TRY(cudaMallocHost(&c.h_views, (size_t)PARAM_RING * HZ_V_COUNT * sizeof(HzView)));
    for(cudaEvent_t& e : c.ring_ev) CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&c.busy, cudaEventDisableTiming));
    if(own_stream)
    {
        CUDA_TRY(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreateWithFlags(&c.done, cudaEventDisableTiming));
    }
    return true;
}

// bytes of device memory one scratch set takes (what ensure_lanes() budgets with)
size_t scratch_bytes(const Slot& s)
{
    const size_t px = s.target_pixels;
    return px * (sizeof(unsigned long long) + 3 + sizeof(float)) +
           ((size_t)s.tri_capacity + (size_t)s.nt * s.nt + (size_t)s.nb * s.nb + 2 * (size_t)s.N) * sizeof(uint32_t) +
           2 * (size_t)s.big_capacity * sizeof(uint2) + (size_t)s.bigtri_capacity * 6 * sizeof(uint4);
}

// Makes sure up to n lanes exist (n <= n_lanes_max) and returns how many there are to use: lanes are only added
// while they fit into half of the device memory that is free right now (a 36000 x 4000 panorama needs ~3 GB per lane).
int ensure_lanes(Slot& s, int n)
{
    if(s.fork_ev == nullptr && cudaEventCreateWithFlags(&s.fork_ev, cudaEventDisableTiming) != cudaSuccess) return 0;
    while((int)s.lanes.size() < n)
    {
        size_t free_b = 0, total_b = 0;
        if(cudaMemGetInfo(&free_b, &total_b) != cudaSuccess || scratch_bytes(s) > free_b / 2) break;
        Scratch c;
        if(!alloc_scratch(s, c, true) || !alloc_scratch_target(s, c, true))
        {
            MSG("Could not allocate render lane %d", (int)s.lanes.size());
            cudaGetLastError();
            free_scratch(c);
            break;
        }
        s.lanes.push_back(c);
    }
    return (int)s.lanes.size() < n ? (int)s.lanes.size() : n;
}

void destroy_slot(Slot* s)
{
    if(s == nullptr) return;
    DeviceGuard g(s->device);
    if(s->stream) cudaStreamSynchronize(s->stream);
    for(Scratch& l : s->lanes) if(l.stream) cudaStreamSynchronize(l.stream);
    free_target(*s);
    free_scratch(s->main);
    for(Scratch& l : s->lanes) free_scratch(l);
    if(s->fork_ev) cudaEventDestroy(s->fork_ev);
    cudaFree(s->d_mosaic);
    for(int i = 0; i < 4; i++) for(int j = 0; j < 4; j++) cudaFree((void*)s->tiles.tile[i][j]);
    for(cudaEvent_t e : s->prof_events) cudaEventDestroy(e);
    cudaFree(s->d_mm_block); cudaFree(s->d_mm_tile);
    if(s->stream) cudaStreamDestroy(s->stream);
    delete s;
}

// vertex.glsl:34-38 on the host, in float (round() = round-half-even, oracle rule F9)
float unwrap_near_rad(float x, float near)
{
    const float d = (x - near) / (2.f * PI_F);
    return (d - rintf(d)) * 2.f * PI_F + near;
}

// lib:1007-1012: tan(elevation) of each GL row as the reference's read-back computes it -- rows of the
// lower half directly, rows of the upper half as the negated value of their mirror row.
void fill_tanel(float* out, int W, int H, float az_deg0, float az_deg1)
{
    const float aspect = (float)W / (float)H;
    for(int y = 0; y < H; y++)
    {
        const int   ysrc   = (y < H / 2 || ((H & 1) && y == H / 2)) ? y : H - 1 - y;
        const float el_ndc = ((float)ysrc + 0.5f) / (float)H * 2.f - 1.f;
        const float el     = el_ndc * (az_deg1 - az_deg0) / 2.f / aspect * M_PI / 180.0f;
        const float t      = tanf(el);
        out[y] = (ysrc != y) ? -t : t;
    }
}

// returns the device pointer of the row table for this window, uploading it if it is not cached
bool tanel_for(Slot& s, float az_deg0, float az_deg1, cudaStream_t st, const float** d_out)
{
    const float daz = az_deg1 - az_deg0;
    for(int k = 0; k < TANEL_SLOTS; k++)
        if(s.tanel_key[k].valid && s.tanel_key[k].W == s.W && s.tanel_key[k].H == s.H &&
           memcmp(&s.tanel_key[k].daz, &daz, sizeof(float)) == 0)
        {
            *d_out = s.d_tanel + (size_t)k * s.H;
            return true;
        }
    const int k = s.tanel_next;
    s.tanel_next = (s.tanel_next + 1) % TANEL_SLOTS;
    // the pinned row may still be in flight from an earlier upload, the device row may still be in use
    CUDA_TRY(cudaStreamSynchronize(st));
    if(st != s.stream) CUDA_TRY(cudaStreamSynchronize(s.stream));
    fill_tanel(s.h_tanel + (size_t)k * s.H, s.W, s.H, az_deg0, az_deg1);
    CUDA_TRY(cudaMemcpyAsync(s.d_tanel + (size_t)k * s.H, s.h_tanel + (size_t)k * s.H,
                             (size_t)s.H * sizeof(float), cudaMemcpyHostToDevice, st));
    s.tanel_key[k] = { true, daz, s.W, s.H };
    *d_out = s.d_tanel + (size_t)k * s.H;
    return true;
}

// the kernels of one render, in order, reading their parameters from sc.d_views; hv = the host copy of those
const Slot::Bands& bands_of(const Slot& s, const Scratch& sc)
{
    return (&sc == &s.main) ? s.bands_single : s.bands_batch;
}

// Zoomed-in views (small angle per pixel) show triangles many pixels large even far from the eye: each band then draws
// its large triangles before the next band is tested against the visibility buffer.  In wide views the far bands have
// hardly any, and one k_big after the last band saves a launch per band.
bool big_after_every_band(const Slot& s, const ViewState& vs)
{
    return fabsf(vs.az_deg1 - vs.az_deg0) < 0.05f * (float)s.W;
}

bool launch_chain(Slot& s, Scratch& sc, const HzView* hv, bool big_per_band, bool worst_case, bool resolve, cudaStream_t st,
                  cudaEvent_t* ev, int* launches)
{
    const HzView* dv = sc.d_views;
    int n = 0;
    if(ev) CUDA_TRY(cudaEventRecord(ev[0], st));
    CUDA_TRY(hz_launch_prepare(hv[HZ_V_NEAR], dv + HZ_V_NEAR, st)); n++;
    if(ev) CUDA_TRY(cudaEventRecord(ev[1], st));
    CUDA_TRY(hz_launch_near(hv[HZ_V_NEAR], dv + HZ_V_NEAR, st)); n++;
    CUDA_TRY(hz_launch_raster(hv[HZ_V_NEAR], dv + HZ_V_NEAR, st)); n++;
    if(ev) CUDA_TRY(cudaEventRecord(ev[2], st));
    CUDA_TRY(hz_launch_big(hv[HZ_V_NEAR], dv + HZ_V_NEAR, st)); n++;
    if(ev) CUDA_TRY(cudaEventRecord(ev[3], st));
    for(int b = 0; b < bands_of(s, sc).n; b++)
    {
        const bool last = (b + 1 == bands_of(s, sc).n);
        int k = 0;
        CUDA_TRY(hz_launch_band(hv[HZ_V_BAND0 + b], dv + HZ_V_BAND0 + b, worst_case, st, &k));
        n += k;
        if(last && ev) CUDA_TRY(cudaEventRecord(ev[4], st));
        // all bands share one queue and counter unless every band has its own k_big (see enqueue_render)
        if(last || (big_per_band && k > 0)) { CUDA_TRY(hz_launch_big(hv[HZ_V_BAND0 + b], dv + HZ_V_BAND0 + b, st)); n++; }
    }
    if(ev) CUDA_TRY(cudaEventRecord(ev[5], st));
    if(resolve) { CUDA_TRY(hz_launch_resolve(hv[HZ_V_NEAR], dv + HZ_V_NEAR, st)); n++; }
    if(ev) CUDA_TRY(cudaEventRecord(ev[6], st));
    *launches = n;
    return true;
}

// Captures the standard chain (full width, vectorised resolve, grids sized for any eye position) once per scratch
// set; every later standard render is one cudaGraphLaunch after the parameter copy.  A dozen separate launches cost
// more host time than the GPU needs for the render.
bool capture_graph(Slot& s, Scratch& sc, const HzView* hv, bool big_per_band)
{
    cudaStream_t cs = nullptr;
    if(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); return false; }
    cudaGraph_t g = nullptr;
    bool ok = cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
    int launches = 0;
    if(ok)
    {
        ok = launch_chain(s, sc, hv, big_per_band, true, true, cs, nullptr, &launches);
        if(cudaStreamEndCapture(cs, &g) != cudaSuccess) ok = false;
    }
    if(ok && cudaGraphInstantiate(&sc.graph, g, 0) != cudaSuccess) { ok = false; sc.graph = nullptr; }
    if(g) cudaGraphDestroy(g);
    cudaStreamDestroy(cs);
    if(!ok)
    {
        cudaGetLastError();
        MSG("CUDA graph capture of the render chain failed; launching the kernels one by one instead");
        return false;
    }
    sc.graph_launches = launches;
    sc.graph_big_per_band = big_per_band;
    return true;
}

// where a render's outputs go (see HzView::n_out): one destination shaped like the target, or the full panoramas
// of several ranks
struct OutSpec
{
    int n = 1;
    uint8_t* image[HZ_MAX_OUT] = {};
    float*   ranges[HZ_MAX_OUT] = {};
    int stride = 0;              // pixels per destination row; 0 = the target's own width (x1-x0)
    int x_off = 0;               // column of the destination where the target's first column goes
};

OutSpec single_out(uint8_t* d_image, float* d_ranges)
{
    OutSpec o;
    o.image[0] = d_image; o.ranges[0] = d_ranges;
    return o;
}

// enqueue one render of columns [x0,x1) on stream st, using scratch set sc, with outputs to `out` (device memory)
bool enqueue_render(Slot& s, Scratch& sc, const ViewState& vs, int x0, int x1, const OutSpec& out, cudaStream_t st)
{
    uint8_t* const d_image = out.image[0];
    float* const d_ranges = out.ranges[0];
    // a scratch set serves one render at a time: if its previous render went to another stream, wait for that one
    if(sc.busy_recorded && sc.last_stream != st) CUDA_TRY(cudaStreamWaitEvent(st, sc.busy, 0));
    HzView v{};
    v.mosaic = s.d_mosaic; v.N = s.N; v.pitch = s.pitch;
    v.e_tab = sc.d_e; v.n_tab = sc.d_n;
    v.mm_block = s.d_mm_block; v.nb = s.nb;
    v.mm_tile  = s.d_mm_tile;  v.nt = s.nt;
    v.viewer_cell_i = vs.viewer_cell_i; v.viewer_cell_j = vs.viewer_cell_j; v.viewer_z = vs.viewer_z;
    v.deg_per_cell = 1.0f / (float)s.cpd;                                // lib:577
    v.cos_viewer_lat = vs.cos_viewer_lat;
    v.curvature = s.curvature;

    // vertex.glsl:139-150, float
    const float az_rad0 = vs.az_deg0 * 0.017453292519943295f;            // radians()
    float       az_rad1 = vs.az_deg1 * 0.017453292519943295f;
    az_rad1 = unwrap_near_rad(az_rad1 - az_rad0, PI_F) + az_rad0;
    v.az_center      = (az_rad0 + az_rad1) / 2.f;
    v.az_ndc_per_rad = 2.0f / (az_rad1 - az_rad0);
    v.aspect         = (float)s.W / (float)s.H;                          // lib:658-659
    v.seam_period    = s.seam_wrap ? v.az_ndc_per_rad * 2.f * PI_F : 0.0f;

    v.znear = s.znear; v.zfar = s.zfar; v.znear_color = s.znear_color; v.zfar_color = s.zfar_color;
    v.W = s.W; v.H = s.H; v.x0 = x0; v.x1 = x1;
    v.vis = sc.d_vis;
    v.stats      = s.collect_stats ? sc.d_counters + STATS_AT : nullptr;
    v.tile_queue = sc.d_tile_queue; v.block_queue = sc.d_block_queue;
    v.tri_queue = sc.d_tri_queue; v.tri_capacity = s.tri_capacity;
    v.bigtri = sc.d_bigtri; v.bigtri_count = sc.d_counters + 3; v.bigtri_capacity = s.bigtri_capacity;
    v.occl_tile_max_pix = s.occl_tile_max_pix; v.occl_block_max_pix = s.occl_block_max_pix;
    v.small_max_pix = s.small_max_pix;
    v.grid_percent = (&sc == &s.main) ? s.grid_percent_single : s.grid_percent_batch;
    v.big_capacity = s.big_capacity;

    // the eye's tile, and how many rings of tiles around it form the foreground pass
    {
        const int ti = (int)floorf(vs.viewer_cell_i / (float)HZ_TILE_CELLS), tj = (int)floorf(vs.viewer_cell_j / (float)HZ_TILE_CELLS);
        v.eye_ti = ti < 0 ? 0 : (ti >= s.nt ? s.nt - 1 : ti);
        v.eye_tj = tj < 0 ? 0 : (tj >= s.nt ? s.nt - 1 : tj);
        v.near_rings = s.near_rings;
    }

    // for the depth bound of hz_rect_test: the diagonal of one cell on the ground
    {
        const double cn = (double)v.deg_per_cell * 6371000.0 * M_PI / 180.0, ce = cn * fabs((double)vs.cos_viewer_lat);
        v.cell_diag2 = (float)((ce * ce + cn * cn) * 1.01);
        // 1/64 pixel for snapping (1/512) and the few-ulp wobble of the angle functions, plus 8 float ulps of the
        // largest window coordinate (an ulp of x at column 36000 is already 1/256 pixel)
        v.box_margin = 0.015625f + 8.0f * 6e-8f * (float)(s.W > s.H ? s.W : s.H);
        v.inv_zrange = (s.zfar > s.znear) ? 1.0f / (s.zfar - s.znear) : 0.0f;   // 0: no far/occlusion culling
    }

    const float* d_tanel = nullptr;
    if(d_ranges && !tanel_for(s, vs.az_deg0, vs.az_deg1, st, &d_tanel)) return false;

    v.counters = sc.d_counters; v.ncounters = N_COUNTERS;
    v.tanel = d_tanel;
    v.n_out = out.n; v.out_stride = out.stride > 0 ? out.stride : x1 - x0; v.out_x0 = out.x_off;
    for(int d = 0; d < out.n; d++) { v.out_image[d] = out.image[d]; v.out_ranges[d] = out.ranges[d]; }

    // ---- the parameter block: variants of v for the near pass, the far queue and each band, into a slot of the
    // pinned ring, then one small copy to the device
    const int slot = sc.ring_next;
    sc.ring_next = (sc.ring_next + 1) % PARAM_RING;
    CUDA_TRY(cudaEventSynchronize(sc.ring_ev[slot]));             // the copy that last used this slot is done
    HzView* hv = sc.h_views + (size_t)slot * HZ_V_COUNT;
    for(int k = 0; k < HZ_V_COUNT; k++) hv[k] = v;
    hv[HZ_V_NEAR].tri_count = sc.d_counters + 2;
    hv[HZ_V_NEAR].big_queue = sc.d_big_queue;                 hv[HZ_V_NEAR].big_count = sc.d_counters + 0;
    for(int k = HZ_V_FAR; k < HZ_V_COUNT; k++)
    {
        hv[k].big_queue = sc.d_big_queue + s.big_capacity;      hv[k].big_count = sc.d_counters + 1;
    }
    const bool big_per_band = big_after_every_band(s, vs);
    {
        int lo = s.near_rings + 1;
        const Slot::Bands& bands = bands_of(s, sc);
        for(int b = 0; b < bands.n; b++)
        {
            HzView& vb = hv[HZ_V_BAND0 + b];
            vb.ring_lo = lo; vb.ring_hi = bands.end[b] > lo ? bands.end[b] : lo;
            vb.tile_count = sc.d_counters + 4 + 4 * b; vb.block_count = sc.d_counters + 5 + 4 * b;
            vb.tri_count  = sc.d_counters + 6 + 4 * b;
            // one k_big per band: each band counts its own entries from 0 (the queue memory is reused, the bands run
            // one after the other); one k_big at the end: all bands append to the same count
            vb.big_count  = sc.d_counters + 7 + (big_per_band ? 4 * b : 0);
            lo = vb.ring_hi;
        }
    }
    CUDA_TRY(cudaMemcpyAsync(sc.d_views, hv, HZ_V_COUNT * sizeof(HzView), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaEventRecord(sc.ring_ev[slot], st));

    // ---- the kernels: a replay of the captured graph where the chain has its standard shape, one by one otherwise
    const bool standard = !s.profiling && s.use_graphs && x0 == 0 && x1 == s.W && (d_image || d_ranges) &&
                          out.n == 1 && v.out_stride == s.W && out.x_off == 0 && hz_resolve_is_vectorisable(v);
    if(standard && !sc.graph_failed)
    {
        if(sc.graph != nullptr && sc.graph_big_per_band != big_per_band)
        {
            // the chain changes shape (rare: the caller went from a wide to a zoomed-in window or back): let the old
            // graph's last launch finish before it is destroyed
            if(sc.busy_recorded) CUDA_TRY(cudaEventSynchronize(sc.busy));
            drop_graph(sc);
        }
        if(sc.graph == nullptr && !capture_graph(s, sc, hv, big_per_band)) sc.graph_failed = true;
        if(sc.graph != nullptr)
        {
            CUDA_TRY(cudaGraphLaunch(sc.graph, st));
            CUDA_TRY(cudaEventRecord(sc.busy, st)); sc.last_stream = st; sc.busy_recorded = true;
            s.launches_last = sc.graph_launches;
            if(&sc == &s.main) s.have_render = true;
            return true;
        }
    }

    cudaEvent_t* ev = nullptr;
    if(s.profiling)
    {
        while(s.prof_events.size() < s.prof_used + PROF_EVENTS)
        {
            cudaEvent_t e;
            CUDA_TRY(cudaEventCreate(&e));
            s.prof_events.push_back(e);
        }
        ev = &s.prof_events[s.prof_used];
        s.prof_used += PROF_EVENTS;
    }
    int launches = 0;
    if(!launch_chain(s, sc, hv, big_per_band, false, d_image || d_ranges, st, ev, &launches)) return false;
    CUDA_TRY(cudaEventRecord(sc.busy, st)); sc.last_stream = st; sc.busy_recorded = true;
    s.launches_last = launches;
    if(&sc == &s.main) s.have_render = (x0 == 0 && x1 == s.W);
    return true;
}

// lib:765-789, float as written there.  The four samples come from the host mmaps (dem.h).
bool compute_move(const horizonator_context_t* ctx, float* viewer_z, float lat, float lon, ViewState& vs)
{
    const horizonator_dem_context_t* d = &ctx->dems;
    const float vci = (lon - (float)d->origin_dem_lon_lat[0]) * (float)d->cells_per_deg - (float)d->origin_dem_cellij[0];
    const float vcj = (lat - (float)d->origin_dem_lon_lat[1]) * (float)d->cells_per_deg - (float)d->origin_dem_cellij[1];
    const int i0 = (int)floorf(vci), j0 = (int)floorf(vcj);
    float z;
    if(viewer_z == nullptr || *viewer_z < 0)
    {
        // a little above the ground so that the bumps right next to the eye don't fill the view
        z = (float)((double)fmaxf(fmaxf((float)horizonator_dem_sample(d, i0,     j0),
                                        (float)horizonator_dem_sample(d, i0 + 1, j0)),
                                  fmaxf((float)horizonator_dem_sample(d, i0,     j0 + 1),
                                        (float)horizonator_dem_sample(d, i0 + 1, j0 + 1))) + 1.0);
        if(viewer_z != nullptr) *viewer_z = z;
    }
    else
        z = *viewer_z;
    vs.viewer_cell_i  = vci;
    vs.viewer_cell_j  = vcj;
    vs.viewer_z       = z;
    vs.cos_viewer_lat = cosf((float)((double)lat * M_PI / (double)180.0f));   // lib:799
    return true;
}

// Device -> caller's host buffer on `st`.  Page-locked destinations (cudaHostAlloc / cudaHostRegister /
// horizonator_host_alloc) are written by DMA at PCIe speed; for ordinary pageable memory the driver stages the
// copy itself, which measured faster here than a private pinned bounce buffer plus memcpy.
bool copy_to_host(void* dst, const void* src, size_t bytes, cudaStream_t st)
{
    CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st));
    return true;
}

} // namespace

extern "C" {

bool horizonator_init(horizonator_context_t* ctx,
                      float viewer_lat, float viewer_lon, float* viewer_z,
                      int offscreen_width, int offscreen_height,
                      int render_radius_cells, float render_radius_m,
                      bool use_glut, bool render_texture, bool SRTM1,
                      const char* dir_dems, const char* dir_tiles,
                      const char* tiles_name, const char* tiles_url_fmt,
                      bool allow_downloads)
{
    (void)dir_tiles; (void)tiles_name; (void)tiles_url_fmt; (void)allow_downloads;
    memset(ctx, 0, sizeof(*ctx));

    if(render_texture)
    {
        MSG("render_texture=true (OpenStreetMap texturing) is not supported by the CUDA renderer");
        return false;
    }
    if(offscreen_width > 0 && offscreen_height <= 0)
    {
        MSG("offscreen_width > 0 needs offscreen_height > 0");
        return false;
    }
    if(dir_dems == nullptr)                                             // lib:94-97
        dir_dems = SRTM1 ? "~/.horizonator/DEMs_SRTM1" : "~/.horizonator/DEMs_SRTM3";

    int ndev = 0;
    if(cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0)
    {
        MSG("No usable CUDA device: libhorizonator renders on the GPU only (there is no CPU path)");
        return false;
    }
    int dev = 0;
    if(const char* env = getenv("HORIZONATOR_DEVICE")) dev = atoi(env);
    else if(cudaGetDevice(&dev) != cudaSuccess) dev = 0;

    if(!horizonator_dem_init(&ctx->dems, viewer_lat, viewer_lon, render_radius_cells, render_radius_m, dir_dems, SRTM1))
    {
        MSG("Couldn't init DEMs. Giving up");
        return false;
    }

    Slot* s = new Slot;
    s->device = dev;
    bool ok = false;
    do
    {
        DeviceGuard guard(dev);
        if(!guard.ok) { MSG("cudaSetDevice(%d) failed", dev); break; }
        auto fail = [](cudaError_t e, const char* what) {
            if(e != cudaSuccess) MSG("CUDA error: %s: %s", what, cudaGetErrorString(e));
            return e != cudaSuccess;
        };
        if(fail(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking), "cudaStreamCreate")) break;

        const horizonator_dem_context_t* d = &ctx->dems;
        s->cpd = d->cells_per_deg;
        s->N   = 2 * d->radius_cells;
        s->pitch = ((s->N + 2) + 63) / 64 * 64;

        // raw tiles to the device (16 spare bytes: k_mosaic's aligned 32-bit loads may touch them)
        s->tiles.cpd = s->cpd;
        s->tiles.origin_cell[0] = d->origin_dem_cellij[0]; s->tiles.origin_cell[1] = d->origin_dem_cellij[1];
        s->tiles.ntiles[0] = d->Ndems_ij[0];               s->tiles.ntiles[1] = d->Ndems_ij[1];
        bool bad = false;
        for(int i = 0; i < d->Ndems_ij[0] && !bad; i++)
            for(int j = 0; j < d->Ndems_ij[1] && !bad; j++)
            {
                if(d->dems[i][j] == nullptr) continue;
                uint8_t* p = nullptr;
                bad = fail(cudaMalloc(&p, d->mmap_sizes[i][j] + 16), "cudaMalloc(tile)") ||
                      fail(cudaMemcpyAsync(p, d->dems[i][j], d->mmap_sizes[i][j], cudaMemcpyHostToDevice, s->stream),
                           "cudaMemcpy(tile)");
                s->tiles.tile[i][j] = p;
            }
        if(bad) break;

        if(fail(cudaMalloc(&s->d_mosaic, (size_t)s->N * s->pitch * sizeof(int16_t)), "cudaMalloc(mosaic)")) break;
        s->nb = (s->N - 1 + HZ_BLOCK_CELLS - 1) / HZ_BLOCK_CELLS;
        s->nt = (s->N - 1 + HZ_TILE_CELLS - 1) / HZ_TILE_CELLS;
        if(const char* env = getenv("HORIZONATOR_NEAR_RINGS")) s->near_rings = atoi(env) < 0 ? 0 : atoi(env);
        if(const char* env = getenv("HORIZONATOR_OCCL_TILE_PIX"))  s->occl_tile_max_pix  = atoi(env);
        if(const char* env = getenv("HORIZONATOR_OCCL_BLOCK_PIX")) s->occl_block_max_pix = atoi(env);
        if(const char* env = getenv("HORIZONATOR_SMALL_PIX"))      s->small_max_pix = atoi(env);
        if(const char* env = getenv("HORIZONATOR_GRID_SCALE"))       s->grid_percent_single = atoi(env);
        if(const char* env = getenv("HORIZONATOR_GRID_SCALE_BATCH")) s->grid_percent_batch  = atoi(env);
        // HORIZONATOR_BANDS / HORIZONATOR_BANDS_BATCH: comma-separated rings at which the bands end (lone views / views
        // of a batch; the first also sets the second unless that is given); the last band always runs to the edge
        auto parse_bands = [](const char* env, Slot::Bands& out) {
            int n = 0;
            for(const char* p = env; *p && n < MAX_BANDS - 1; )
            {
                const int r = atoi(p);
                if(r > 0) out.end[n++] = r;
                while(*p && *p != ',') p++;
                if(*p == ',') p++;
            }
            out.end[n++] = 1 << 20;
            out.n = n;
        };
        if(const char* env = getenv("HORIZONATOR_BANDS")) { parse_bands(env, s->bands_single); s->bands_batch = s->bands_single; }
        if(const char* env = getenv("HORIZONATOR_BANDS_BATCH")) parse_bands(env, s->bands_batch);
        if(const char* env = getenv("HORIZONATOR_GRAPHS")) s->use_graphs = atoi(env) != 0;
        if(const char* env = getenv("HORIZONATOR_LANES")) s->n_lanes_max = atoi(env) < 1 ? 1 : (atoi(env) > 32 ? 32 : atoi(env));
        if(fail(cudaMalloc(&s->d_mm_block, (size_t)s->nb * s->nb * sizeof(short2)), "cudaMalloc(pyramid)")) break;
        if(fail(cudaMalloc(&s->d_mm_tile, (size_t)s->nt * s->nt * sizeof(short2)), "cudaMalloc(pyramid)")) break;
        if(!alloc_scratch(*s, s->main, false)) break;
        if(fail(hz_launch_mosaic(s->tiles, s->d_mosaic, s->N, s->pitch, s->stream), "k_mosaic")) break;
        if(fail(hz_launch_pyramid(s->d_mosaic, s->N, s->pitch, s->d_mm_block, s->nb, s->d_mm_tile, s->nt, s->stream),
                "k_minmax")) break;

        // without an offscreen size the reference opens a 1024x1024 window (lib:142)
        const int W = offscreen_width > 0 ? offscreen_width : 1024;
        const int H = offscreen_width > 0 ? offscreen_height : 1024;
        if(!alloc_target(*s, W, H)) break;
        if(fail(cudaStreamSynchronize(s->stream), "DEM upload/decode")) break;
        ok = true;
    } while(0);

    if(!ok)
    {
        destroy_slot(s);
        horizonator_dem_deinit(&ctx->dems);
        memset(ctx, 0, sizeof(*ctx));
        return false;
    }

    {
        std::lock_guard<std::mutex> lock(g_table_mutex);
        size_t k = 0;
        while(k < g_table.size() && g_table[k] != nullptr) k++;
        if(k == g_table.size()) g_table.push_back(s); else g_table[k] = s;
        ctx->program = (uint32_t)(k + 1);
    }

    const int n1 = 2 * ctx->dems.radius_cells - 1;
    ctx->Ntriangles     = n1 * n1 * 2;                                  // lib:202-203
    ctx->render_texture = false;
    ctx->use_glut       = use_glut;
    ctx->glut_window    = 1;
    // the 17 GL uniform locations of the reference: no meaning here, -1 = "no such uniform"
    memset(&ctx->uniform_aspect, 0xFF,
           offsetof(horizonator_context_t, uniform_zfar_color) + sizeof(int32_t) - offsetof(horizonator_context_t, uniform_aspect));
    if(offscreen_width > 0)
    {
        ctx->offscreen.inited = true;
        ctx->offscreen.width  = offscreen_width;
        ctx->offscreen.height = offscreen_height;
    }

    horizonator_move(ctx, viewer_z, viewer_lat, viewer_lon);           // lib:611
    horizonator_set_zextents(ctx, HORIZONATOR_ZNEAR_DEFAULT, HORIZONATOR_ZFAR_DEFAULT,
                             HORIZONATOR_ZNEAR_DEFAULT, HORIZONATOR_ZFAR_DEFAULT);
    horizonator_pan_zoom(ctx, -45.f, 45.f);                             // lib:670
    return true;
}

void horizonator_deinit(horizonator_context_t* ctx)
{
    if(ctx == nullptr) return;
    Slot* s = nullptr;
    if(ctx->Ntriangles > 0 && ctx->program != 0)
    {
        std::lock_guard<std::mutex> lock(g_table_mutex);
        if(ctx->program <= g_table.size())
        {
            s = g_table[ctx->program - 1];
            g_table[ctx->program - 1] = nullptr;
        }
    }
    destroy_slot(s);
    if(ctx->Ntriangles > 0) horizonator_dem_deinit(&ctx->dems);
    ctx->Ntriangles  = 0;
    ctx->program     = 0;
    ctx->glut_window = 0;
    ctx->offscreen.inited = false;
}

bool horizonator_resized(const horizonator_context_t* ctx, int width, int height)
{
    Slot* s = slot_of(ctx);
    if(s == nullptr) return false;
    if(ctx->offscreen.inited)
    {
        MSG("Resizing an offscreen context is not supported");          // the reference asserts here
        return false;
    }
    if(width <= 0 || height <= 0) return false;
    DeviceGuard g(s->device);
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    return alloc_target(*s, width, height);
}

bool horizonator_pan_zoom(const horizonator_context_t* ctx, float az_deg0, float az_deg1)
{
    Slot* s = slot_of(ctx);
    if(s == nullptr) return false;
    s->view.az_deg0 = az_deg0;      // stored as given, like the uniforms at lib:833-834
    s->view.az_deg1 = az_deg1;
    return true;
}

bool horizonator_move(horizonator_context_t* ctx, float* viewer_z, float viewer_lat, float viewer_lon)
{
    Slot* s = slot_of(ctx);
    if(s == nullptr) return false;
    const float az0 = s->view.az_deg0, az1 = s->view.az_deg1;
    if(!compute_move(ctx, viewer_z, viewer_lat, viewer_lon, s->view)) return false;
    s->view.az_deg0 = az0; s->view.az_deg1 = az1;
    ctx->viewer_lat = viewer_lat;                                       // lib:812-813
    ctx->viewer_lon = viewer_lon;
    return true;
}

bool horizonator_set_zextents(horizonator_context_t* ctx,
                              float znear, float zfar, float znear_color, float zfar_color)
{
    Slot* s = slot_of(ctx);
    if(s == nullptr) return false;
    if(!(znear > 0.0f && znear_color > 0.0f && zfar > 0.0f && zfar_color > 0.0f)) return false;   // lib:875-877
    s->znear = znear; s->zfar = zfar; s->znear_color = znear_color; s->zfar_color = zfar_color;
    return true;
}

bool horizonator_redraw(const horizonator_context_t* ctx)
{
    Slot* s = slot_of(ctx);
    if(s == nullptr) return false;
    DeviceGuard g(s->device);
    if(!enqueue_render(*s, s->main, s->view, 0, s->W, single_out(s->d_image, s->d_ranges), s->stream)) return false;
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    return true;
}

bool horizonator_render_offscreen(const horizonator_context_t* ctx, char* image, float* ranges)
{
    Slot* s = slot_of(ctx);
    if(s == nullptr) return false;
    if(!ctx->offscreen.inited)
    {
        MSG("Prior to calling horizonator_render_offscreen(), the context must have been inited for offscreen rendering with horizonator_init(offscreen_width,height > 0)");
        return false;
    }
    DeviceGuard g(s->device);
    const size_t px = (size_t)s->W * s->H;
    if(!enqueue_render(*s, s->main, s->view, 0, s->W,
                       single_out(image ? s->d_image : nullptr, ranges ? s->d_ranges : nullptr), s->stream)) return false;
    if(image  && !copy_to_host(image,  s->d_image,  px * 3, s->stream)) return false;
    if(ranges && !copy_to_host(ranges, s->d_ranges, px * sizeof(float), s->stream)) return false;
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    return true;
}

bool horizonator_pick(const horizonator_context_t* ctx, float* lat, float* lon, int x, int y)
{
    Slot* s = slot_of(ctx);
    if(s == nullptr || !s->have_render) return false;
    if(x < 0 || y < 0 || x >= s->W || y >= s->H) return false;
    DeviceGuard g(s->device);
    unsigned long long key = 0;
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    CUDA_TRY(cudaMemcpy(&key, s->main.d_vis + (size_t)(s->H - 1 - y) * s->W + x, sizeof(key), cudaMemcpyDeviceToHost));
    const float depth = (float)((double)(unsigned)(key >> 40) * (1.0 / 16777215.0));
    if(depth >= 1.0f) return false;                                     // lib:1272
    // lib:1282-1295: the depth is treated as horizontal distance
    const double range_en = depth * (s->zfar - s->znear) + s->znear;
    return horizonator_unproject(lat, lon, x, y, -1., range_en,
                                 ctx->viewer_lat, s->view.cos_viewer_lat, ctx->viewer_lon,
                                 s->view.az_deg0, s->view.az_deg1, s->W, s->H);
}

// ---- pure host geometry (lib:1053-1213): double precision, no device involved ----------------------------

static double unwrap_near_rad_d(double x, double near)
{
    const double d = (x - near) / (2. * M_PI);
    return (d - round(d)) * 2. * M_PI + near;
}

bool horizonator_x_from_az(double* x, double* az_ndc_per_rad,
                           double az_rad, double az_rad0, double az_rad1, int width)
{
    az_rad1 = unwrap_near_rad_d(az_rad1 - az_rad0, M_PI) + az_rad0;
    const double center = (az_rad0 + az_rad1) / 2.;
    az_rad = unwrap_near_rad_d(az_rad, center);
    const double per_rad = 2.0 / (az_rad1 - az_rad0);
    const double az_ndc  = (az_rad - center) * per_rad;
    if(!(-1. <= az_ndc && az_ndc <= 1.)) return false;
    if(az_ndc_per_rad != nullptr) *az_ndc_per_rad = per_rad;
    *x = (az_ndc + 1.) / 2. * width - 0.5;                               // NDC [-1,1] -> pixel (-0.5, W-0.5)
    return true;
}

bool horizonator_project(double* x, double* y, double* range,
                         double lat_viewer, double cos_lat_viewer, double lon_viewer, double ele_viewer,
                         double lat, double lon, double ele,
                         double az_rad0, double az_rad1, int width, int height)
{
    const float Rearth = 6371000.0;
    const double dlat = (lat - lat_viewer) * M_PI / 180;
    const double dlon = (lon - lon_viewer) * M_PI / 180;
    const double east  = dlon * Rearth * cos_lat_viewer;
    const double north = dlat * Rearth;
    const double d2 = east * east + north * north;

    double per_rad;
    if(!horizonator_x_from_az(x, &per_rad, atan2(east, north), az_rad0, az_rad1, width)) return false;

    const double h = ele - ele_viewer;
    *range = sqrt(d2 + h * h);
    const double aspect = (double)width / (double)height;
    const double el_ndc = atan2(h, sqrt(d2)) * aspect * per_rad;
    if(!(-1. <= el_ndc && el_ndc <= 1.)) return false;
    *y = (-el_ndc + 1.) / 2. * height - 0.5;
    return true;
}

bool horizonator_unproject(float* lat, float* lon, int x, int y,
                           double range_enh, double range_en,
                           double lat_viewer, double cos_lat_viewer, double lon_viewer,
                           double az_deg0, double az_deg1, int width, int height)
{
    if(1 != (range_enh > 0.) + (range_en > 0.)) return false;
    const float Rearth = 6371000.0;
    // mixed float/double exactly as lib:1185-1186
    const float az_ndc = ((float)x + 0.5f) / (float)width * 2.f - 1.f;
    const float az     = (az_ndc * (az_deg1 - az_deg0) / 2.f + (az_deg1 + az_deg0) / 2.f) * M_PI / 180.0f;
    if(range_en <= 0)
    {
        const double aspect = (double)width / (double)height;
        const double el_ndc = ((double)y + 0.5) / (double)height * 2. - 1.;
        const double el     = el_ndc * (az_deg1 - az_deg0) / 2. / aspect * M_PI / 180.0;
        range_en = cos(el) * range_enh;
    }
    const float e = range_en * sinf(az);
    const float n = range_en * cosf(az);
    *lon = lon_viewer + e / Rearth / M_PI * 180. / cos_lat_viewer;
    *lat = lat_viewer + n / Rearth / M_PI * 180.;
    return true;
}

// ---- additive API (include/horizonator-batch.h) -----------------------------------------------------------

// Common part of the two batch calls.  The views are dealt round-robin to render lanes, each with its own stream
// and scratch, so that the (latency-bound) kernel chains of different views overlap; the lanes start after
// everything already queued on `st` and `st` continues after all of them.  to_host: outputs go through the lane's
// device staging buffers and a device->host copy on the lane's stream, which overlaps the next views' kernels.
static bool render_batch_common(const horizonator_context_t* ctx, Slot* s, int n, const horizonator_view_t* views,
                                uint8_t* images, float* ranges, bool to_host, cudaStream_t st)
{
    const size_t px = (size_t)s->W * s->H;
    std::vector<ViewState> vs((size_t)n);
    for(int k = 0; k < n; k++)
    {
        float z = views[k].viewer_z;
        if(!compute_move(ctx, &z, views[k].lat, views[k].lon, vs[k])) return false;
        vs[k].az_deg0 = views[k].az_deg0; vs[k].az_deg1 = views[k].az_deg1;
    }

    // Lanes need every per-window row table of the batch resident before they start (tanel_for() synchronises
    // when it has to upload one): resolve them all on `st`, then check that none evicted another.
    int n_lanes = n < s->n_lanes_max ? n : s->n_lanes_max;
    if(n_lanes > 1 && ranges != nullptr)
    {
        const float* dummy;
        for(int k = 0; k < n; k++) if(!tanel_for(*s, vs[k].az_deg0, vs[k].az_deg1, st, &dummy)) return false;
        const int before = s->tanel_next;
        for(int k = 0; k < n; k++) if(!tanel_for(*s, vs[k].az_deg0, vs[k].az_deg1, st, &dummy)) return false;
        if(s->tanel_next != before) n_lanes = 1;      // more distinct windows than table slots: one at a time
    }
    if(n_lanes > 1) n_lanes = ensure_lanes(*s, n_lanes);      // as many as fit; fewer than 2: one view at a time

    if(n_lanes <= 1)
    {
        for(int k = 0; k < n; k++)
        {
            uint8_t* di = images ? (to_host ? s->d_image  : images + (size_t)k * px * 3) : nullptr;
            float*   dr = ranges ? (to_host ? s->d_ranges : ranges + (size_t)k * px)     : nullptr;
            if(!enqueue_render(*s, s->main, vs[k], 0, s->W, single_out(di, dr), st)) return false;
            if(to_host)
            {
                if(images && !copy_to_host(images + (size_t)k * px * 3, di, px * 3, st)) return false;
                if(ranges && !copy_to_host(ranges + (size_t)k * px, dr, px * sizeof(float), st)) return false;
            }
        }
        s->have_render = false;     // the visibility buffer no longer matches the context's own view
        return true;
    }

    CUDA_TRY(cudaEventRecord(s->fork_ev, st));
    for(int l = 0; l < n_lanes; l++) CUDA_TRY(cudaStreamWaitEvent(s->lanes[l].stream, s->fork_ev, 0));
    for(int k = 0; k < n; k++)
    {
        Scratch& lane = s->lanes[k % n_lanes];
        uint8_t* di = images ? (to_host ? lane.d_image  : images + (size_t)k * px * 3) : nullptr;
        float*   dr = ranges ? (to_host ? lane.d_ranges : ranges + (size_t)k * px)     : nullptr;
        if(!enqueue_render(*s, lane, vs[k], 0, s->W, single_out(di, dr), lane.stream)) return false;
        if(to_host)
        {
            if(images && !copy_to_host(images + (size_t)k * px * 3, di, px * 3, lane.stream)) return false;
            if(ranges && !copy_to_host(ranges + (size_t)k * px, dr, px * sizeof(float), lane.stream)) return false;
        }
    }
    for(int l = 0; l < n_lanes; l++)
    {
        CUDA_TRY(cudaEventRecord(s->lanes[l].done, s->lanes[l].stream));
        CUDA_TRY(cudaStreamWaitEvent(st, s->lanes[l].done, 0));
    }
    return true;
}

bool horizonator_render_batch_device(const horizonator_context_t* ctx, int n, const horizonator_view_t* views,
                                     void* d_images, void* d_ranges, void* stream)
{
    Slot* s = slot_of(ctx);
    if(s == nullptr || n < 0 || (n > 0 && views == nullptr)) return false;
    DeviceGuard g(s->device);
    cudaStream_t st = stream ? (cudaStream_t)stream : s->stream;
    if(!render_batch_common(ctx, s, n, views, (uint8_t*)d_images, (float*)d_ranges, false, st)) return false;
    if(stream == nullptr) CUDA_TRY(cudaStreamSynchronize(st));
    return true;
}

bool horizonator_render_batch(const horizonator_context_t* ctx, int n, const horizonator_view_t* views,
                              char* images, float* ranges)
{
    Slot* s = slot_of(ctx);
    if(s == nullptr || n < 0 || (n > 0 && views == nullptr)) return false;
    DeviceGuard g(s->device);
    if(!render_batch_common(ctx, s, n, views, (uint8_t*)images, ranges, true, s->stream)) return false;
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    return true;
}

bool horizonator_render_wedge_device(const horizonator_context_t* ctx, int x0, int x1,
                                     void* d_image, void* d_ranges, void* stream)
{
    Slot* s = slot_of(ctx);
    if(s == nullptr) return false;
    if(x0 < 0 || x1 > s->W || x0 >= x1)
    {
        MSG("wedge columns [%d,%d) are not inside [0,%d)", x0, x1, s->W);
        return false;
    }
    DeviceGuard g(s->device);
    cudaStream_t st = stream ? (cudaStream_t)stream : s->stream;
    if(!enqueue_render(*s, s->main, s->view, x0, x1, single_out((uint8_t*)d_image, (float*)d_ranges), st)) return false;
    if(stream == nullptr) CUDA_TRY(cudaStreamSynchronize(st));
    return true;
}

// ---- wedge-sharded panoramas assembled over NVLink ---------------------------------------------------------

bool horizonator_peer_alloc(const horizonator_context_t* ctx, size_t bytes, void** d_ptr, unsigned char handle[64])
{
    Slot* s = slot_of(ctx);
    if(s == nullptr || d_ptr == nullptr || handle == nullptr || bytes == 0) return false;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    DeviceGuard g(s->device);
    void* p = nullptr;
    CUDA_TRY(cudaMalloc(&p, bytes));
    cudaIpcMemHandle_t h;
    const cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if(e != cudaSuccess)
    {
        MSG("cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
        cudaFree(p);
        return false;
    }
    memcpy(handle, &h, 64);
    *d_ptr = p;
    return true;
}

bool horizonator_peer_open(const horizonator_context_t* ctx, const unsigned char handle[64], void** d_ptr)
{
    Slot* s = slot_of(ctx);
    if(s == nullptr || d_ptr == nullptr || handle == nullptr) return false;
    DeviceGuard g(s->device);
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    CUDA_TRY(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return true;
}

bool horizonator_peer_close(const horizonator_context_t* ctx, void* d_ptr)
{
    Slot* s = slot_of(ctx);
    if(s == nullptr || d_ptr == nullptr) return false;
    DeviceGuard g(s->device);
    CUDA_TRY(cudaIpcCloseMemHandle(d_ptr));
    return true;
}

bool horizonator_peer_free(const horizonator_context_t* ctx, void* d_ptr)
{
    Slot* s = slot_of(ctx);
    if(s == nullptr) return false;
    DeviceGuard g(s->device);
    CUDA_TRY(cudaFree(d_ptr));
    return true;
}

bool horizonator_render_wedge_peers(const horizonator_context_t* ctx, int x0, int x1, int n_peers,
                                    void* const* d_images, void* const* d_ranges, void* stream)
{
    Slot* s = slot_of(ctx);
    if(s == nullptr) return false;
    if(x0 < 0 || x1 > s->W || x0 >= x1)
    {
        MSG("wedge columns [%d,%d) are not inside [0,%d)", x0, x1, s->W);
        return false;
    }
    if(n_peers < 1 || n_peers > HZ_MAX_OUT || (d_images == nullptr && d_ranges == nullptr))
    {
        MSG("need 1..%d destinations and at least one kind of output", HZ_MAX_OUT);
        return false;
    }
    OutSpec out;
    out.n = n_peers; out.stride = s->W; out.x_off = x0;
    for(int d = 0; d < n_peers; d++)
    {
        out.image[d]  = d_images ? (uint8_t*)d_images[d] : nullptr;
        out.ranges[d] = d_ranges ? (float*)d_ranges[d]   : nullptr;
        if((d_images && out.image[d] == nullptr) || (d_ranges && out.ranges[d] == nullptr))
        {
            MSG("destination %d is null", d);
            return false;
        }
    }
    DeviceGuard g(s->device);
    cudaStream_t st = stream ? (cudaStream_t)stream : s->stream;
    if(!enqueue_render(*s, s->main, s->view, x0, x1, out, st)) return false;
    if(stream == nullptr) CUDA_TRY(cudaStreamSynchronize(st));
    return true;
}

bool horizonator_peer_barrier(const horizonator_context_t* ctx, int n_ranks, int rank, void* const* d_flags,
                              unsigned int epoch, void* stream)
{
    Slot* s = slot_of(ctx);
    if(s == nullptr || d_flags == nullptr || n_ranks < 1 || n_ranks > HZ_MAX_OUT || rank < 0 || rank >= n_ranks) return false;
    HzPeerFlags f{};
    f.n = n_ranks; f.rank = rank;
    for(int r = 0; r < n_ranks; r++)
    {
        if(d_flags[r] == nullptr) return false;
        f.arrive[r] = (uint32_t*)d_flags[r];
    }
    DeviceGuard g(s->device);
    cudaStream_t st = stream ? (cudaStream_t)stream : s->stream;
    CUDA_TRY(hz_launch_peer_barrier(f, epoch, st));
    if(stream == nullptr) CUDA_TRY(cudaStreamSynchronize(st));
    return true;
}

bool horizonator_set_seam_wrap(const horizonator_context_t* ctx, bool on)
{
    Slot* s = slot_of(ctx);
    if(s == nullptr) return false;
    s->seam_wrap = on;
    return true;
}

bool horizonator_set_earth_curvature(const horizonator_context_t* ctx, bool on, float refraction)
{
    Slot* s = slot_of(ctx);
    if(s == nullptr) return false;
    if(on && !(refraction >= 0.f && refraction < 1.f))
    {
        MSG("refraction coefficient %g is not in [0,1)", (double)refraction);
        return false;
    }
    s->curvature = on ? (1.0f - refraction) / (2.0f * 6371000.0f) : 0.0f;
    return true;
}

void* horizonator_host_alloc(size_t bytes)
{
    void* p = nullptr;
    if(cudaMallocHost(&p, bytes) != cudaSuccess)
    {
        MSG("cudaMallocHost(%zu) failed", bytes);
        cudaGetLastError();
        return nullptr;
    }
    return p;
}

void horizonator_host_free(void* p)
{
    if(p != nullptr) cudaFreeHost(p);
}

bool horizonator_download_mosaic(const horizonator_context_t* ctx, int16_t* mosaic)
{
    Slot* s = slot_of(ctx);
    if(s == nullptr || mosaic == nullptr) return false;
    DeviceGuard g(s->device);
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    CUDA_TRY(cudaMemcpy2D(mosaic, (size_t)s->N * sizeof(int16_t), s->d_mosaic, (size_t)s->pitch * sizeof(int16_t),
                          (size_t)s->N * sizeof(int16_t), s->N, cudaMemcpyDeviceToHost));
    return true;
}

bool horizonator_time_mosaic(const horizonator_context_t* ctx, int reps, float* ms_per_run)
{
    Slot* s = slot_of(ctx);
    if(s == nullptr || reps <= 0 || ms_per_run == nullptr) return false;
    DeviceGuard g(s->device);
    cudaEvent_t a, b;
    CUDA_TRY(cudaEventCreate(&a));
    CUDA_TRY(cudaEventCreate(&b));
    CUDA_TRY(hz_launch_mosaic(s->tiles, s->d_mosaic, s->N, s->pitch, s->stream));      // warm-up
    CUDA_TRY(cudaEventRecord(a, s->stream));
    for(int k = 0; k < reps; k++) CUDA_TRY(hz_launch_mosaic(s->tiles, s->d_mosaic, s->N, s->pitch, s->stream));
    CUDA_TRY(cudaEventRecord(b, s->stream));
    CUDA_TRY(cudaEventSynchronize(b));
    float ms = 0;
    CUDA_TRY(cudaEventElapsedTime(&ms, a, b));
    cudaEventDestroy(a); cudaEventDestroy(b);
    *ms_per_run = ms / (float)reps;
    return true;
}

bool horizonator_profile_enable(const horizonator_context_t* ctx, bool on)
{
    Slot* s = slot_of(ctx);
    if(s == nullptr) return false;
    s->profiling = on;
    s->collect_stats = on;
    return true;
}

bool horizonator_profile_read(const horizonator_context_t* ctx, float out_ms[6], int* renders)
{
    Slot* s = slot_of(ctx);
    if(s == nullptr || out_ms == nullptr || renders == nullptr) return false;
    DeviceGuard g(s->device);
    double sum[PROF_EVENTS - 1] = {};
    const size_t n = s->prof_used / PROF_EVENTS;
    for(size_t r = 0; r < n; r++)
    {
        cudaEvent_t* ev = &s->prof_events[PROF_EVENTS * r];
        CUDA_TRY(cudaEventSynchronize(ev[PROF_EVENTS - 1]));
        for(int k = 0; k < PROF_EVENTS - 1; k++)
        {
            float ms = 0;
            CUDA_TRY(cudaEventElapsedTime(&ms, ev[k], ev[k + 1]));
            sum[k] += ms;
        }
    }
    for(int k = 0; k < PROF_EVENTS - 1; k++) out_ms[k] = n ? (float)(sum[k] / (double)n) : 0.f;
    *renders = (int)n;
    s->prof_used = 0;
    return true;
}

bool horizonator_last_render_stats(const horizonator_context_t* ctx, unsigned int out[5])
{
    Slot* s = slot_of(ctx);
    if(s == nullptr) return false;
    DeviceGuard g(s->device);
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    unsigned int counters[N_COUNTERS] = {};
    CUDA_TRY(cudaMemcpy(counters, s->main.d_counters, sizeof(counters), cudaMemcpyDeviceToHost));
    out[0] = counters[0];
    for(int b = 0; b < MAX_BANDS; b++) out[0] += counters[7 + 4 * b];
    out[1] = 2 * s->big_capacity; out[2] = s->launches_last; out[3] = (unsigned)s->device;
    out[4] = counters[2];
    for(int b = 0; b < MAX_BANDS; b++) out[4] += counters[6 + 4 * b];     // triangle lists of the near pass and the bands
    return true;
}

bool horizonator_render_counters(const horizonator_context_t* ctx, unsigned int out[16])
{
    Slot* s = slot_of(ctx);
    if(s == nullptr || out == nullptr) return false;
    DeviceGuard g(s->device);
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    CUDA_TRY(cudaMemcpy(out, s->main.d_counters + STATS_AT, HZ_STAT_COUNT * sizeof(unsigned int), cudaMemcpyDeviceToHost));
    return true;
}

bool horizonator_horizon_profile_device(const horizonator_context_t* ctx, const void* d_ranges, int n,
                                        void* d_rows, void* d_range, void* stream)
{
    Slot* s = slot_of(ctx);
    if(s == nullptr || n < 0 || d_ranges == nullptr || d_rows == nullptr || d_range == nullptr) return false;
    DeviceGuard g(s->device);
    cudaStream_t st = stream ? (cudaStream_t)stream : s->stream;
    CUDA_TRY(hz_launch_horizon((const float*)d_ranges, n, s->W, s->H, (int*)d_rows, (float*)d_range, st));
    if(stream == nullptr) CUDA_TRY(cudaStreamSynchronize(st));
    return true;
}

} // extern "C"

/* horizonator.h -- C ABI of libhorizonator, B200-native build (no OpenGL anywhere).
 *
 * Every declaration below is binary compatible with the reference interface it replaces,
 * /root/reference/horizonator.h (struct :13-54, entry points :84-214), so existing callers
 * (the reference's horizonator-pywrap.c, standalone.c, annotator.c, horizonator.cc) link
 * against this library unchanged.  sizeof(horizonator_context_t) == 472; field offsets are
 * listed beside each group and are verified by tests/test_abi.py.
 *
 * What differs behind the ABI:
 *   - the fields that held GL uniform locations / GL object names are kept only for layout.
 *     `program` carries the 32-bit handle of the device-side render state (0 = none); the
 *     other GL-era fields stay 0.
 *   - the terrain is decoded once into HBM and every render runs hand-written sm_100a CUDA
 *     kernels (project+cull+rasterise with a 64-bit atomicMin visibility buffer, resolve).
 *   - nothing aborts: every failure is a `false` return plus a MSG() line on stderr.
 *   - the library fails (returns false) when no CUDA device is usable; there is no CPU path.
 */
#pragma once

#include <stdbool.h>
#include <stdint.h>

#include "dem.h"

#ifdef __cplusplus
extern "C" {
#endif

/* default clip distances in metres (reference horizonator.h:9-10) */
#define HORIZONATOR_ZNEAR_DEFAULT 100.0f
#define HORIZONATOR_ZFAR_DEFAULT  40000.0f

typedef struct
{
    int  Ntriangles;               /* @0   2*(2R-1)^2 once initialised; >0 <=> context valid */
    bool render_texture, use_glut; /* @4,5 as passed to horizonator_init()                    */
    int  glut_window;              /* @8   kept 1 while the context is live (callers of the
                                           reference treat 0 as "closed")                    */

    /* @12..@76: 17 x int32, GL uniform locations in the reference. Unused here (all -1). */
    int32_t uniform_aspect, uniform_az_deg0, uniform_az_deg1;
    int32_t uniform_viewer_cell_i;
    int32_t uniform_viewer_cell_j;
    int32_t uniform_viewer_z;
    int32_t uniform_viewer_lat;
    int32_t uniform_cos_viewer_lat;
    int32_t uniform_texturemap_lon0;
    int32_t uniform_texturemap_lon1;
    int32_t uniform_texturemap_dlat0;
    int32_t uniform_texturemap_dlat1;
    int32_t uniform_texturemap_dlat2;
    int32_t uniform_znear, uniform_zfar;
    int32_t uniform_znear_color, uniform_zfar_color;

    uint32_t program;              /* @80  handle of the device render state                 */

    float viewer_lat, viewer_lon;  /* @84,88 last position given to init()/move()            */

    horizonator_dem_context_t dems;/* @96  truthful: callers read it (slippymap-annotations) */

    struct
    {
        bool     inited;           /* @448 */
        uint32_t frameBufID;       /* @452 unused, 0 */
        uint32_t renderBufID;      /* @456 unused, 0 */
        uint32_t depthBufID;       /* @460 unused, 0 */
        int      width, height;    /* @464,468 */
    } offscreen;
} horizonator_context_t;

__attribute__((unused))
static bool horizonator_context_isvalid(const horizonator_context_t* ctx)
{
    return ctx->Ntriangles > 0;
}

/* Replaces horizonator-lib.c:61-680.  Loads the DEM square around the viewer into HBM and
 * prepares the render state; initial azimuth window is -45..45 deg, initial z extents are
 * the defaults above.  offscreen_width>0 selects an image of that size; otherwise the
 * context renders into an internal 1024x1024 buffer until horizonator_resized() is called.
 * use_glut is recorded but has no effect (there is no window system).  render_texture=true
 * is refused (needs network tile downloads; out of scope) with a `false` return.
 * viewer_z: NULL or *viewer_z<0 => eye = highest of the 4 surrounding samples + 1 m,
 * reported back through the pointer when it is not NULL. */
bool horizonator_init(horizonator_context_t* ctx,
                      float viewer_lat, float viewer_lon,
                      float* viewer_z,
                      int offscreen_width, int offscreen_height,
                      int   render_radius_cells,
                      float render_radius_m,
                      bool use_glut,
                      bool render_texture,
                      bool SRTM1,
                      const char* dir_dems,
                      const char* dir_tiles,
                      const char* tiles_name,
                      const char* tiles_url_fmt,
                      bool allow_downloads);

/* Replaces horizonator-lib.c:682-689.  Frees device and host state (the reference leaks
 * both).  Safe on a zeroed context and idempotent. */
void horizonator_deinit(horizonator_context_t* ctx);

/* Replaces horizonator-lib.c:838-856.  Refused (false) on an offscreen context, where the
 * reference asserts. */
bool horizonator_resized(const horizonator_context_t* ctx, int width, int height);

/* Replaces horizonator-lib.c:818-836.  az_deg0 sits at the left edge of pixel column 0,
 * az_deg1 at the right edge of the last column; elevation scale follows to keep pixels
 * square in angle. */
bool horizonator_pan_zoom(const horizonator_context_t* ctx, float az_deg0, float az_deg1);

/* Replaces horizonator-lib.c:691-816.  Moves the eye inside the already loaded square. */
bool horizonator_move(horizonator_context_t* ctx,
                      float* viewer_z,
                      float viewer_lat, float viewer_lon);

/* Replaces horizonator-lib.c:864-885.  All four must be > 0, otherwise nothing changes and
 * the call returns false (this is what the reference code does, whatever its comment says).
 * The clip planes act on the slant range; the colour ramp on the horizontal distance. */
bool horizonator_set_zextents(horizonator_context_t* ctx,
                              float znear,       float zfar,
                              float znear_color, float zfar_color);

/* Replaces horizonator-lib.c:887-899.  Renders into the device-side buffers only. */
bool horizonator_redraw(const horizonator_context_t* ctx);

/* Replaces horizonator-lib.c:1216-1296.  Uses the depth kept from the last render. */
bool horizonator_pick(const horizonator_context_t* ctx,
                      float* lat, float* lon,
                      int x, int y);

/* Replaces horizonator-lib.c:911-1051.  image: W*H*3 bytes, B,G,R per pixel; ranges: W*H
 * floats; both caller-owned host memory, top row first, either may be NULL.  Pixels that
 * show no terrain: image (255,0,0), range -1. */
bool horizonator_render_offscreen(const horizonator_context_t* ctx,
                                  char* image, float* ranges);

/* Replaces horizonator-lib.c:1062-1095. */
bool horizonator_x_from_az(double* x,
                           double* az_ndc_per_rad,
                           double az_rad,
                           double az_rad0,
                           double az_rad1,
                           int width);

/* Replaces horizonator-lib.c:1097-1155. */
bool horizonator_project(double* x,
                         double* y,
                         double* range,
                         double lat_viewer, double cos_lat_viewer,
                         double lon_viewer,
                         double ele_viewer,
                         double lat,
                         double lon,
                         double ele,
                         double az_rad0,
                         double az_rad1,
                         int width,
                         int height);

/* Replaces horizonator-lib.c:1157-1213.  Exactly one of range_enh / range_en is > 0. */
bool horizonator_unproject(float* lat, float* lon,
                           int x, int y,
                           double range_enh,
                           double range_en,
                           double lat_viewer, double cos_lat_viewer,
                           double lon_viewer,
                           double az_deg0,
                           double az_deg1,
                           int width,
                           int height);

#ifdef __cplusplus
}
#endif

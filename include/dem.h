/* dem.h -- virtual SRTM mosaic, C ABI of libhorizonator (B200-native build).
 *
 * ABI-compatible replacement for the reference interface /root/reference/dem.h:10-66.
 * Layout (x86-64, gcc): sizeof == 352; dems@0 mmap_sizes@128 mmap_fd@256
 * origin_dem_lon_lat@320 origin_dem_cellij@328 Ndems_ij@336 radius_cells@344
 * cells_per_deg@348.  tests/test_abi.py checks these numbers.
 *
 * The mosaic is the square of (2*radius_cells)^2 SRTM samples around the viewer, addressed
 * (i east, j north) from its south-west corner, stitched from at most 4x4 one-degree tiles.
 * The host keeps the tiles mmap'd (callers and horizonator_move() sample single cells from
 * them); the renderer keeps a decoded int16 copy of the whole square in HBM.
 */
#pragma once

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define max_Ndems_ij 4 /* tiles per axis, at most (reference dem.h:8) */

typedef struct
{
    /* [i_lon][j_lat]; NULL = tile absent or empty => elevation 0 */
    unsigned char* dems      [max_Ndems_ij][max_Ndems_ij];
    size_t         mmap_sizes[max_Ndems_ij][max_Ndems_ij];
    int            mmap_fd   [max_Ndems_ij][max_Ndems_ij];

    int origin_dem_lon_lat[2]; /* integer lon,lat naming the tile that holds the SW corner */
    int origin_dem_cellij [2]; /* cell of the SW corner inside that tile                    */
    int Ndems_ij          [2]; /* tiles used along lon, lat                                  */

    int radius_cells;          /* R: the mosaic is 2R x 2R samples                           */
    int cells_per_deg;         /* 1200 (SRTM3) or 3600 (SRTM1)                               */
} horizonator_dem_context_t;

/* Replaces dem.c:78-243.  Exactly one of render_radius_cells / render_radius_m must be > 0.
 * Returns false (after a MSG() on stderr) on: both/neither radius, more than 4 tiles on an
 * axis, a tile of the wrong size, an unusable "~/" path.  A missing or zero-length tile is
 * not an error: it reads as elevation 0. */
bool horizonator_dem_init(horizonator_dem_context_t* ctx,
                          float viewer_lat,
                          float viewer_lon,
                          int   render_radius_cells,
                          float render_radius_m,
                          const char* datadir,
                          bool  SRTM1);

/* Replaces dem.c:245-261.  Safe on a zeroed context, safe to call twice. */
void horizonator_dem_deinit(horizonator_dem_context_t* ctx);

/* Replaces dem.c:264-309.  (i,j) relative to the SW corner; -1 outside the loaded tiles,
 * negative/void samples clamp to 0.  The reference reads out of bounds for i==0 (or j==0)
 * when origin_dem_cellij is 0; here that case reads column (row) 0 of tile 0. */
int16_t horizonator_dem_sample(const horizonator_dem_context_t* ctx, int i, int j);

/* Replaces dem.c:313-330.  Inclusive lat/lon of the first and last cell. */
void horizonator_dem_bounds_latlon_deg(const horizonator_dem_context_t* ctx,
                                       float* lat0, float* lon0,
                                       float* lat1, float* lon1);

#ifdef __cplusplus
}
#endif

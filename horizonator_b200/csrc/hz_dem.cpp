// hz_dem.cpp -- host side of the virtual SRTM mosaic: the four entry points of include/dem.h.
//
// Replaces /root/reference/dem.c.  The tiles stay mmap'd on the host (callers sample single cells, and
// horizonator_move() needs the four samples around the eye); hz_api.cpp uploads the same bytes to the GPU
// where k_mosaic decodes the whole square once.
//
// Float/integer types follow dem.c expression by expression (it computes in float through <tgmath.h>), because
// the tile selection must come out identical: see the line references.
#include "dem.h"
#include "util.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

namespace {

// dem.c:22-76: "<dir>/N34W118.hgt"; a leading "~/" is replaced by $HOME
bool tile_filename(char* path, size_t bufsize, int lat, int lon, const char* datadir)
{
    const char ns = lat >= 0 ? 'N' : 'S';
    const char we = lon >= 0 ? 'E' : 'W';
    int n;
    if(datadir[0] == '~' && datadir[1] == '/')
    {
        const char* home = getenv("HOME");
        if(home == nullptr)
        {
            MSG("User asked for ~, but the 'HOME' env var isn't defined");
            return false;
        }
        n = snprintf(path, bufsize, "%s/%s/%c%.2d%c%.3d.hgt", home, datadir + 2, ns, abs(lat), we, abs(lon));
    }
    else
        n = snprintf(path, bufsize, "%s/%c%.2d%c%.3d.hgt", datadir, ns, abs(lat), we, abs(lon));
    return n >= 0 && (size_t)n < bufsize;
}

} // namespace

extern "C" {

bool horizonator_dem_init(horizonator_dem_context_t* ctx,
                          float viewer_lat, float viewer_lon,
                          int render_radius_cells, float render_radius_m,
                          const char* datadir, bool SRTM1)
{
    if(render_radius_cells < 0 && render_radius_m < 0)
    {
        MSG("Exactly one of (render_radius_cells,render_radius_m) should be >0. Both were <0");
        return false;
    }
    if(render_radius_cells > 0 && render_radius_m > 0)
    {
        MSG("Exactly one of (render_radius_cells,render_radius_m) should be >0. Both were >0");
        return false;
    }

    memset(ctx, 0, sizeof(*ctx));
    const int cpd = SRTM1 ? 3600 : 1200;                       // dem.c:101-104
    ctx->cells_per_deg = cpd;

    if(render_radius_cells > 0)
        ctx->radius_cells = render_radius_cells;
    else
    {
        // dem.c:124-126, in double: the square must contain a circle of render_radius_m, and the
        // east-west cell size shrinks with cos(lat)
        const double Rearth = 6371000.0;
        const double coslat = cos(M_PI / 180.0 * (double)viewer_lat);
        ctx->radius_cells = (int)(0.5 + (double)render_radius_m / (Rearth * M_PI / 180. * coslat / (double)cpd));
    }
    if(ctx->radius_cells <= 0)
    {
        MSG("The render radius came out as %d cells; need at least 1", ctx->radius_cells);
        return false;
    }
    const int R = ctx->radius_cells;

    const float viewer[2] = { viewer_lon, viewer_lat };
    for(int a = 0; a < 2; a++)                                  // dem.c:136-179
    {
        const int   cell0  = (int)floorf(viewer[a] * (float)cpd) - (R - 1);
        const float origin = (float)cell0 / (float)cpd;
        ctx->origin_dem_lon_lat[a] = (int)floorf(origin);
        ctx->origin_dem_cellij[a]  = (int)roundf((origin - (float)ctx->origin_dem_lon_lat[a]) * (float)cpd);

        const int last_cell = ctx->origin_dem_cellij[a] + 2 * R - 1;
        const int last_tile = last_cell / cpd;
        ctx->Ndems_ij[a] = last_tile + 1;
        if(last_cell == last_tile * cpd)
            ctx->Ndems_ij[a]--;          // that cell is also the last one of the previous tile
        if(ctx->Ndems_ij[a] > max_Ndems_ij)
        {
            MSG("Requested radius too large. Increase the compile-time-constant max_Ndems_ij from the current value of %d",
                max_Ndems_ij);
            memset(ctx, 0, sizeof(*ctx));
            return false;
        }
    }

    const off_t expected = (off_t)(cpd + 1) * (cpd + 1) * 2;    // dem.c:129-132
    for(int j = 0; j < ctx->Ndems_ij[1]; j++)                   // dem.c:183-240
        for(int i = 0; i < ctx->Ndems_ij[0]; i++)
        {
            char filename[1024];
            if(!tile_filename(filename, sizeof(filename),
                              j + ctx->origin_dem_lon_lat[1], i + ctx->origin_dem_lon_lat[0], datadir))
            {
                horizonator_dem_deinit(ctx);
                MSG("Couldn't construct DEM filename");
                return false;
            }
            const int fd = open(filename, O_RDONLY);
            if(fd < 0)
            {
                MSG("Warning: couldn't open DEM file '%s'. Assuming elevation=0 (sea surface?)", filename);
                continue;
            }
            struct stat sb;
            if(fstat(fd, &sb) != 0 || sb.st_size == 0)
            {
                close(fd);      // empty file: sea, silently
                continue;
            }
            if(sb.st_size != expected)
            {
                close(fd);
                horizonator_dem_deinit(ctx);
                MSG("The DEM file '%s' has unexpected size. Is this a %d-arc-sec SRTM DEM?", filename, SRTM1 ? 1 : 3);
                return false;
            }
            void* p = mmap(nullptr, sb.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
            if(p == MAP_FAILED)
            {
                close(fd);
                horizonator_dem_deinit(ctx);
                MSG("Couldn't mmap the DEM file '%s'", filename);
                return false;
            }
            ctx->dems[i][j]       = (unsigned char*)p;
            ctx->mmap_sizes[i][j] = (size_t)sb.st_size;
            ctx->mmap_fd[i][j]    = fd;
        }
    return true;
}

void horizonator_dem_deinit(horizonator_dem_context_t* ctx)
{
    for(int i = 0; i < max_Ndems_ij; i++)
        for(int j = 0; j < max_Ndems_ij; j++)
        {
            if(ctx->dems[i][j] != nullptr && ctx->dems[i][j] != (unsigned char*)MAP_FAILED)
                munmap(ctx->dems[i][j], ctx->mmap_sizes[i][j]);
            ctx->dems[i][j] = nullptr;
            ctx->mmap_sizes[i][j] = 0;
            if(ctx->mmap_fd[i][j] > 0) close(ctx->mmap_fd[i][j]);
            ctx->mmap_fd[i][j] = 0;
        }
}

int16_t horizonator_dem_sample(const horizonator_dem_context_t* ctx, int i, int j)
{
    if(i < 0 || j < 0) return -1;                               // dem.c:270
    const int cpd = ctx->cells_per_deg;
    int tile[2], cell[2] = { i + ctx->origin_dem_cellij[0], j + ctx->origin_dem_cellij[1] };
    for(int a = 0; a < 2; a++)
    {
        tile[a]  = cell[a] / cpd;
        cell[a] -= tile[a] * cpd;
        if(cell[a] == 0) { tile[a]--; cell[a] = cpd; }          // dem.c:287-291: shared edge
        if(tile[a] >= ctx->Ndems_ij[a]) return -1;              // dem.c:293
        if(tile[a] < 0) { tile[a] = 0; cell[a] = 0; }           // the reference reads out of bounds here
    }
    const unsigned char* dem = ctx->dems[tile[0]][tile[1]];
    if(dem == nullptr) return 0;
    const size_t p = (size_t)cell[0] + (size_t)(cpd - cell[1]) * (size_t)(cpd + 1);   // north row first
    const int16_t z = (int16_t)((dem[2 * p] << 8) | dem[2 * p + 1]);                  // big-endian
    return z < 0 ? 0 : z;
}

void horizonator_dem_bounds_latlon_deg(const horizonator_dem_context_t* ctx,
                                       float* lat0, float* lon0, float* lat1, float* lon1)
{
    const float cpd = (float)ctx->cells_per_deg;
    const int   n   = 2 * ctx->radius_cells - 1;
    *lon0 = (float)ctx->origin_dem_lon_lat[0] + (float)ctx->origin_dem_cellij[0] / cpd;
    *lat0 = (float)ctx->origin_dem_lon_lat[1] + (float)ctx->origin_dem_cellij[1] / cpd;
    *lon1 = (float)ctx->origin_dem_lon_lat[0] + ((float)ctx->origin_dem_cellij[0] + n) / cpd;
    *lat1 = (float)ctx->origin_dem_lon_lat[1] + ((float)ctx->origin_dem_cellij[1] + n) / cpd;
}

} // extern "C"

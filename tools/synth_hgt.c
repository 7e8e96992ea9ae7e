/* synth_hgt.c -- seeded synthetic SRTM tiles for tests and benchmarks (no network, no real
 * DEMs in this environment).
 *
 * Elevation is a pure function of the GLOBAL cell coordinate (lon*cpd + column, lat*cpd + row)
 * and the seed, so the row/column that neighbouring tiles share (SRTM tiles overlap by one
 * sample) is identical in both files, like real SRTM.  Terrain = ridged multi-octave value
 * noise, roughly 0..3000 m with some "sea" below 0 (the renderer clamps negatives to 0) and
 * optional sparse -32768 voids.
 *
 * File format written = what dem.c of the reference reads: (cpd+1)^2 big-endian int16, first
 * row is the NORTH edge, first column the WEST edge; name N34W118.hgt = SW corner.
 *
 * Build: gcc -O2 -fopenmp -shared -fPIC tools/synth_hgt.c -o tools/libsynth.so -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

static inline uint32_t hash3(uint32_t x, uint32_t y, uint32_t s)
{
    uint32_t h = x * 0x9E3779B1u ^ y * 0x85EBCA77u ^ s * 0xC2B2AE3Du;
    h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12; h *= 0x297A2D39u; h ^= h >> 15;
    return h;
}
static inline float lattice(int32_t x, int32_t y, uint32_t s)
{
    return (float)(hash3((uint32_t)x, (uint32_t)y, s) >> 8) * (1.0f / 16777216.0f);
}
static inline float fade(float t) { return t * t * t * (t * (t * 6.f - 15.f) + 10.f); }

static float value_noise(double x, double y, uint32_t s)
{
    double fx = floor(x), fy = floor(y);
    int32_t ix = (int32_t)fx, iy = (int32_t)fy;
    float tx = fade((float)(x - fx)), ty = fade((float)(y - fy));
    float a = lattice(ix, iy, s),     b = lattice(ix + 1, iy, s);
    float c = lattice(ix, iy + 1, s), d = lattice(ix + 1, iy + 1, s);
    float ab = a + (b - a) * tx, cd = c + (d - c) * tx;
    return ab + (cd - ab) * ty;
}

/* elevation in metres at global cell (gx, gy) of a grid with cpd cells per degree */
static int16_t elevation(int64_t gx, int64_t gy, int cpd, uint32_t seed, int voids)
{
    const double lon = (double)gx / cpd, lat = (double)gy / cpd;
    double freq = 3.0;     /* first octave: features of ~1/3 degree */
    float amp = 1.0f, sum = 0.0f, norm = 0.0f;
    for(int o = 0; o < 9; o++)
    {
        float n = value_noise(lon * freq + 1000.0, lat * freq + 1000.0, seed + 101u * o);
        if(o >= 2) n = 1.0f - fabsf(2.0f * n - 1.0f);      /* ridged upper octaves */
        sum += amp * n; norm += amp;
        amp *= 0.5f; freq *= 2.0;
    }
    float f = sum / norm;                                   /* 0..1 */
    float z = 3600.0f * powf(f, 1.6f) - 500.0f;
    if(voids && (hash3((uint32_t)gx, (uint32_t)gy, seed ^ 0xABCDu) % 20011u) == 0) return -32768;
    if(z > 32000.f) z = 32000.f;
    return (int16_t)lrintf(z);
}

/* Fill buf[(cpd+1)^2] (big-endian, north row first) for the tile whose SW corner is (lat, lon). */
void synth_tile(uint8_t* buf, int lat, int lon, int cpd, uint32_t seed, int voids)
{
    const int n = cpd + 1;
    #pragma omp parallel for schedule(dynamic, 16)
    for(int r = 0; r < n; r++)
    {
        const int64_t gy = (int64_t)lat * cpd + (cpd - r);
        for(int c = 0; c < n; c++)
        {
            const int64_t gx = (int64_t)lon * cpd + c;
            uint16_t z = (uint16_t)elevation(gx, gy, cpd, seed, voids);
            buf[2 * ((size_t)r * n + c) + 0] = (uint8_t)(z >> 8);
            buf[2 * ((size_t)r * n + c) + 1] = (uint8_t)(z & 0xFF);
        }
    }
}

/* Write <dir>/N34W118.hgt style file. Returns 0 on success. */
int synth_write_tile(const char* dir, int lat, int lon, int cpd, uint32_t seed, int voids)
{
    const size_t n = (size_t)(cpd + 1) * (cpd + 1) * 2;
    uint8_t* buf = (uint8_t*)malloc(n);
    if(!buf) return -1;
    synth_tile(buf, lat, lon, cpd, seed, voids);
    char path[1024];
    snprintf(path, sizeof(path), "%s/%c%02d%c%03d.hgt", dir,
             lat >= 0 ? 'N' : 'S', abs(lat), lon >= 0 ? 'E' : 'W', abs(lon));
    FILE* f = fopen(path, "wb");
    if(!f) { free(buf); return -2; }
    size_t w = fwrite(buf, 1, n, f);
    fclose(f); free(buf);
    return w == n ? 0 : -3;
}

/* horizonator-standalone -- GL-free command-line renderer on top of libhorizonator (B200-native build).
 *
 * The offscreen half of the reference's `standalone` tool (/root/reference/standalone.c:113-512), with the same
 * options and conventions where they apply:
 *
 *   horizonator-standalone --width W [--height H] --image OUT.png [--ranges OUT.f32] [--SRTM1]
 *       [--znear M] [--zfar M] [--znear-color M] [--zfar-color M] [--dirdems DIR] [--cut-off-bottom-px N]
 *       [--pois POIS.csv --labels OUT.json]
 *       LAT LON AZ_CENTER_DEG AZ_RADIUS_DEG
 *
 *   - AZ_CENTER/AZ_RADIUS refer to the CENTRES of the first and last pixel column; the viewport is half a pixel wider
 *     on each side (standalone.c:403-404)
 *   - without --height the image gets a 20-degree vertical field of view (standalone.c:406-411)
 *   - the DEM radius is --zfar metres, the eye sits 1 m above the terrain (standalone.c:433-442)
 *   - --znear-color/--zfar-color default to --znear/--zfar (standalone.c:333-334)
 *
 * What the reference's tool does through FreeImage, GLUT and cairo is replaced or left out: the PNG is written by the
 * small encoder below (stored deflate blocks, no compression library needed); there is no window mode; instead of an
 * annotated PDF/SVG the tool can emit the annotation GEOMETRY as JSON: which points of interest are visible and
 * where their markers go, following annotator.c:280-348 (projection with horizonator_project(), 500 m .. 100 km,
 * vertical search of +-6 pixels in the range image for the nearest match within 500 m).  --ranges dumps the range
 * image as raw little-endian float32, row-major, top row first (-1 = no terrain).
 */
#define _GNU_SOURCE
#include <float.h>
#include <getopt.h>
#include <math.h>
#include <stdbool.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>

#include "horizonator.h"
#include "util.h"

/* ------------------------------------------------------------------------------------------------ PNG */

static uint32_t crc_table[256];
static void crc_init(void)
{
    for(uint32_t n = 0; n < 256; n++)
    {
        uint32_t c = n;
        for(int k = 0; k < 8; k++) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
        crc_table[n] = c;
    }
}
static uint32_t crc_update(uint32_t c, const uint8_t* p, size_t n)
{
    while(n--) c = crc_table[(c ^ *p++) & 0xFF] ^ (c >> 8);
    return c;
}
static void put32(uint8_t* p, uint32_t v) { p[0] = v >> 24; p[1] = v >> 16; p[2] = v >> 8; p[3] = v; }

static bool png_chunk(FILE* f, const char type[4], const uint8_t* data, uint32_t n)
{
    uint8_t head[8], tail[4];
    put32(head, n); memcpy(head + 4, type, 4);
    uint32_t c = crc_update(0xFFFFFFFFu, head + 4, 4);
    if(n) c = crc_update(c, data, n);
    put32(tail, c ^ 0xFFFFFFFFu);
    return fwrite(head, 1, 8, f) == 8 && (n == 0 || fwrite(data, 1, n, f) == n) && fwrite(tail, 1, 4, f) == 4;
}

/* 8-bit RGB PNG from rows of B,G,R bytes; the zlib stream uses stored (uncompressed) deflate blocks */
static bool write_png_bgr(const char* path, const uint8_t* bgr, int width, int height)
{
    const size_t row = (size_t)width * 3 + 1, raw = row * (size_t)height;
    const size_t nblocks = (raw + 65534) / 65535;
    const size_t zlen = 2 + raw + 5 * (nblocks ? nblocks : 1) + 4;
    uint8_t* z = malloc(zlen);
    uint8_t* scan = malloc(raw ? raw : 1);
    if(!z || !scan) { free(z); free(scan); return false; }

    for(int y = 0; y < height; y++)
    {
        uint8_t* o = scan + (size_t)y * row;
        const uint8_t* in = bgr + (size_t)y * width * 3;
        *o++ = 0;                                           /* filter type: none */
        for(int x = 0; x < width; x++, in += 3) { *o++ = in[2]; *o++ = in[1]; *o++ = in[0]; }
    }
    size_t zp = 0;
    z[zp++] = 0x78; z[zp++] = 0x01;
    uint32_t a = 1, b = 0;                                  /* Adler-32 */
    size_t done = 0;
    do
    {
        const size_t n = raw - done > 65535 ? 65535 : raw - done;
        z[zp++] = (done + n == raw) ? 1 : 0;                /* BFINAL, BTYPE = 00 */
        z[zp++] = n & 0xFF; z[zp++] = n >> 8; z[zp++] = ~n & 0xFF; z[zp++] = (~n >> 8) & 0xFF;
        memcpy(z + zp, scan + done, n);
        for(size_t k = 0; k < n; k++) { a += scan[done + k]; if(a >= 65521) a -= 65521; b += a; if(b >= 65521) b -= 65521; }
        zp += n; done += n;
    } while(done < raw);
    put32(z + zp, (b << 16) | a); zp += 4;

    FILE* f = fopen(path, "wb");
    bool ok = f != NULL;
    if(ok)
    {
        static const uint8_t sig[8] = { 0x89, 'P', 'N', 'G', '\r', '\n', 0x1A, '\n' };
        uint8_t ihdr[13];
        put32(ihdr, (uint32_t)width); put32(ihdr + 4, (uint32_t)height);
        ihdr[8] = 8; ihdr[9] = 2; ihdr[10] = 0; ihdr[11] = 0; ihdr[12] = 0;
        crc_init();
        ok = fwrite(sig, 1, 8, f) == 8 && png_chunk(f, "IHDR", ihdr, 13) && png_chunk(f, "IDAT", z, (uint32_t)zp) &&
             png_chunk(f, "IEND", NULL, 0);
        ok = (fclose(f) == 0) && ok;
    }
    free(z); free(scan);
    return ok;
}

/* ------------------------------------------------------------------------------------------------ annotation geometry */

#define MARKER_DIST_MIN   500.0        /* annotator.c:19-20 */
#define MARKER_DIST_MAX   100000.0
#define MARKER_RANGE_FUZZ 500.0        /* annotator.c:22-23 */
#define MARKER_PIXEL_FUZZ 6

typedef struct { char name[128]; double lat, lon, ele; } poi_t;

static int read_pois(const char* path, poi_t** out)
{
    FILE* f = fopen(path, "r");
    if(!f) return -1;
    int n = 0, cap = 64;
    poi_t* p = malloc(cap * sizeof(*p));
    char line[512];
    while(p && fgets(line, sizeof(line), f))
    {
        if(line[0] == '#' || line[0] == '\n') continue;
        poi_t q;
        /* name,lat,lon,elevation_m */
        char* c1 = strchr(line, ',');
        if(!c1) continue;
        const size_t len = (size_t)(c1 - line) < sizeof(q.name) - 1 ? (size_t)(c1 - line) : sizeof(q.name) - 1;
        memcpy(q.name, line, len); q.name[len] = 0;
        if(sscanf(c1 + 1, "%lf,%lf,%lf", &q.lat, &q.lon, &q.ele) != 3) continue;
        if(n == cap) { cap *= 2; p = realloc(p, cap * sizeof(*p)); if(!p) break; }
        p[n++] = q;
    }
    fclose(f);
    *out = p;
    return p ? n : -1;
}

/* annotator.c:280-348: which POIs are seen, and at which pixel their marker goes */
static bool write_labels(const char* path, const poi_t* pois, int npois, const float* ranges, int width, int height_out,
                         int height, double lat, double lon, double eye_z, double az_deg0, double az_deg1)
{
    FILE* f = fopen(path, "w");
    if(!f) return false;
    const double coslat = cos(lat * M_PI / 180.);
    fprintf(f, "[");
    int nout = 0;
    for(int i = 0; i < npois; i++)
    {
        double x, y, range;
        if(!horizonator_project(&x, &y, &range, lat, coslat, lon, eye_z, pois[i].lat, pois[i].lon, pois[i].ele,
                                az_deg0 * M_PI / 180., az_deg1 * M_PI / 180., width, height)) continue;
        if(range < MARKER_DIST_MIN || range > MARKER_DIST_MAX) continue;
        const long col = lround(x);
        if(col < 0 || col >= width) continue;
        /* scan down the column around the predicted row for the rendered range closest to the POI's */
        int best = 0;
        double best_err = DBL_MAX;
        for(int d = -MARKER_PIXEL_FUZZ; d < MARKER_PIXEL_FUZZ; d++)
        {
            if(y + (double)d < 0) continue;
            if(y + (double)d >= height_out) break;
            const long r = lround(y) + d;
            if(r < 0 || r >= height_out) continue;
            const float have = ranges[(size_t)width * (size_t)r + (size_t)col];
            if(have <= 0.0f) continue;
            const double err = fabs(range - (double)have);
            if(err < best_err) { best_err = err; best = d; }
            else break;                                     /* ranges only get nearer further down */
        }
        if(best_err >= MARKER_RANGE_FUZZ) continue;         /* hidden behind something else */
        fprintf(f, "%s\n {\"name\": \"", nout++ ? "," : "");
        for(const char* c = pois[i].name; *c; c++)
        {
            if(*c == '"' || *c == '\\') fputc('\\', f);
            if((unsigned char)*c >= 0x20) fputc(*c, f);
        }
        fprintf(f, "\", \"x\": %.3f, \"y\": %.3f, \"range_m\": %.1f, \"range_rendered_m\": %.1f}", x, y + (double)best, range,
                (double)ranges[(size_t)width * (size_t)(lround(y) + best) + (size_t)col]);
    }
    fprintf(f, "\n]\n");
    return fclose(f) == 0;
}

/* ------------------------------------------------------------------------------------------------ main */

static void usage(const char* argv0)
{
    fprintf(stderr,
            "%s --width WIDTH_PIXELS [--height HEIGHT_PIXELS] --image OUT.png [--ranges OUT.f32]\n"
            "   [--SRTM1] [--znear M] [--zfar M] [--znear-color M] [--zfar-color M]\n"
            "   [--dirdems DIRECTORY] [--cut-off-bottom-px N] [--pois POIS.csv --labels OUT.json]\n"
            "   LAT LON AZ_CENTER_DEG AZ_RADIUS_DEG\n\n"
            "Renders the terrain seen from LAT,LON into a PNG (red = far, black = near, blue = sky) on a CUDA device.\n"
            "AZ_..._DEG refer to the centres of the first and last pixel columns.  Without --height a 20-degree\n"
            "vertical field of view is used.  DEMs are read from --dirdems or ~/.horizonator/DEMs_SRTM3 (DEMs_SRTM1).\n"
            "There is no window mode, no --texture and no PDF/SVG output in this build.\n", argv0);
}

int main(int argc, char* argv[])
{
    enum { OPT_RANGES = 1000, OPT_POIS, OPT_LABELS, OPT_ZNEAR, OPT_ZFAR, OPT_ZNEARC, OPT_ZFARC, OPT_DEMS, OPT_TEXTURE };
    static const struct option opts[] = {
        { "width", required_argument, NULL, 'w' }, { "height", required_argument, NULL, 'H' },
        { "cut-off-bottom-px", required_argument, NULL, 'c' }, { "image", required_argument, NULL, 'i' },
        { "ranges", required_argument, NULL, OPT_RANGES }, { "SRTM1", no_argument, NULL, 'S' },
        { "znear", required_argument, NULL, OPT_ZNEAR }, { "zfar", required_argument, NULL, OPT_ZFAR },
        { "znear-color", required_argument, NULL, OPT_ZNEARC }, { "zfar-color", required_argument, NULL, OPT_ZFARC },
        { "dirdems", required_argument, NULL, OPT_DEMS }, { "pois", required_argument, NULL, OPT_POIS },
        { "labels", required_argument, NULL, OPT_LABELS }, { "texture", no_argument, NULL, OPT_TEXTURE },
        { "help", no_argument, NULL, 'h' }, { NULL, 0, NULL, 0 } };

    int width = 0, height = 0, cut = 0;
    bool srtm1 = false;
    const char *image_path = NULL, *ranges_path = NULL, *pois_path = NULL, *labels_path = NULL, *dir_dems = NULL;
    float znear = HORIZONATOR_ZNEAR_DEFAULT, zfar = HORIZONATOR_ZFAR_DEFAULT, znear_color = -1.f, zfar_color = -1.f;

    int o;
    /* '+': stop at the first positional argument, so that negative longitudes are not taken for options */
    while((o = getopt_long(argc, argv, "+h", opts, NULL)) != -1)
        switch(o)
        {
        case 'w': width = atoi(optarg); break;
        case 'H': height = atoi(optarg); break;
        case 'c': cut = atoi(optarg); break;
        case 'i': image_path = optarg; break;
        case 'S': srtm1 = true; break;
        case OPT_RANGES: ranges_path = optarg; break;
        case OPT_POIS: pois_path = optarg; break;
        case OPT_LABELS: labels_path = optarg; break;
        case OPT_ZNEAR: znear = (float)atof(optarg); break;
        case OPT_ZFAR: zfar = (float)atof(optarg); break;
        case OPT_ZNEARC: znear_color = (float)atof(optarg); break;
        case OPT_ZFARC: zfar_color = (float)atof(optarg); break;
        case OPT_DEMS: dir_dems = optarg; break;
        case OPT_TEXTURE: fprintf(stderr, "--texture is not supported by the CUDA renderer\n"); return 1;
        case 'h': usage(argv[0]); return 0;
        default: usage(argv[0]); return 1;
        }
    if(argc - optind != 4)
    {
        fprintf(stderr, "Need exactly 4 non-option arguments. Got %d\n\n", argc - optind);
        usage(argv[0]);
        return 1;
    }
    if(width <= 0 || image_path == NULL)
    {
        fprintf(stderr, "--width and --image are required (there is no window mode in this build)\n\n");
        usage(argv[0]);
        return 1;
    }
    const size_t ilen = strlen(image_path);
    if(ilen < 5 || strcasecmp(image_path + ilen - 4, ".png") != 0)
    {
        fprintf(stderr, "--image MUST be given a '.png' filename (annotated .pdf/.svg output is not part of this build;\n"
                        "use --pois/--labels for the annotation geometry)\n");
        return 1;
    }
    if((pois_path == NULL) != (labels_path == NULL))
    {
        fprintf(stderr, "--pois and --labels go together\n");
        return 1;
    }
    if(znear_color < 0.f) znear_color = znear;
    if(zfar_color  < 0.f) zfar_color  = zfar;

    const float lat = (float)atof(argv[optind + 0]), lon = (float)atof(argv[optind + 1]);
    const float az_center = (float)atof(argv[optind + 2]);
    float az_radius = (float)atof(argv[optind + 3]);
    if(lat < -80.f || lat > 80.f)   { fprintf(stderr, "Got invalid latitude\n");  return 1; }
    if(lon < -180.f || lon > 180.f) { fprintf(stderr, "Got invalid longitude\n"); return 1; }
    if(width < 2) { fprintf(stderr, "--width must be at least 2\n"); return 1; }

    /* pixel-centre convention: the viewport is half a pixel wider on each side */
    const float az_per_pixel = (float)(2. * az_radius / (float)(width - 1));
    az_radius += az_per_pixel / 2.f;
    if(height <= 0) height = (int)roundf((float)width * 20.0f / az_radius);
    if(cut < 0 || cut >= height) { fprintf(stderr, "--cut-off-bottom-px out of range\n"); return 1; }

    uint8_t* image = malloc((size_t)width * height * 3);
    float* ranges = malloc((size_t)width * height * sizeof(float));
    if(!image || !ranges) { MSG("image,ranges buffer malloc() failed"); return 1; }

    horizonator_context_t ctx;
    float viewer_z = -1.0f;
    if(!horizonator_init(&ctx, lat, lon, &viewer_z, width, height, -1, zfar, true, false, srtm1,
                         dir_dems, NULL, NULL, NULL, false))
    {
        fprintf(stderr, "horizonator_init() failed\n");
        return 1;
    }
    int rc = 1;
    if(!horizonator_set_zextents(&ctx, znear, zfar, znear_color, zfar_color))
        fprintf(stderr, "horizonator_set_zextents() failed\n");
    else if(!horizonator_pan_zoom(&ctx, az_center - az_radius, az_center + az_radius))
        fprintf(stderr, "horizonator_pan_zoom() failed\n");
    else if(!horizonator_render_offscreen(&ctx, (char*)image, ranges))
        fprintf(stderr, "render failed\n");
    else if(!write_png_bgr(image_path, image, width, height - cut))
        fprintf(stderr, "Couldn't save to '%s'\n", image_path);
    else
    {
        rc = 0;
        if(ranges_path)
        {
            FILE* f = fopen(ranges_path, "wb");
            const size_t n = (size_t)width * (size_t)(height - cut);
            if(!f || fwrite(ranges, sizeof(float), n, f) != n) { fprintf(stderr, "Couldn't save to '%s'\n", ranges_path); rc = 1; }
            if(f) fclose(f);
        }
        if(pois_path)
        {
            poi_t* pois = NULL;
            const int n = read_pois(pois_path, &pois);
            if(n < 0) { fprintf(stderr, "Couldn't read '%s'\n", pois_path); rc = 1; }
            else if(!write_labels(labels_path, pois, n, ranges, width, height - cut, height, lat, lon, viewer_z,
                                  az_center - az_radius, az_center + az_radius))
            {
                fprintf(stderr, "Couldn't save to '%s'\n", labels_path);
                rc = 1;
            }
            free(pois);
        }
        fprintf(stderr, "Rendered %dx%d from %.6f,%.6f (eye at %.1f m) az %.3f..%.3f\n", width, height, lat, lon, viewer_z,
                az_center - az_radius, az_center + az_radius);
    }
    horizonator_deinit(&ctx);
    free(image); free(ranges);
    return rc;
}

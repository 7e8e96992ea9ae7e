/* freeimage_stubs.c -- TEST INFRASTRUCTURE (oracle).  The reference's texture branch (out of scope, never
 * reached with render_texture=false) is the only user of FreeImage; these satisfy the linker and report
 * "no such image" if anything does call them. */
#include <FreeImage.h>
#include <stddef.h>

FREE_IMAGE_FORMAT FreeImage_GetFileType(const char* filename, int size) { (void)filename; (void)size; return FIF_UNKNOWN; }
FIBITMAP* FreeImage_Load(FREE_IMAGE_FORMAT fif, const char* filename, int flags) { (void)fif; (void)filename; (void)flags; return NULL; }
FREE_IMAGE_COLOR_TYPE FreeImage_GetColorType(FIBITMAP* dib) { (void)dib; return 0; }
FIBITMAP* FreeImage_ConvertTo24Bits(FIBITMAP* dib) { (void)dib; return NULL; }
void      FreeImage_Unload(FIBITMAP* dib) { (void)dib; }
unsigned  FreeImage_GetWidth(FIBITMAP* dib)  { (void)dib; return 0; }
unsigned  FreeImage_GetHeight(FIBITMAP* dib) { (void)dib; return 0; }
unsigned  FreeImage_GetBPP(FIBITMAP* dib)    { (void)dib; return 0; }
unsigned  FreeImage_GetPitch(FIBITMAP* dib)  { (void)dib; return 0; }
BYTE*     FreeImage_GetBits(FIBITMAP* dib)   { (void)dib; return NULL; }

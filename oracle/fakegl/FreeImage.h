/* FreeImage.h -- TEST INFRASTRUCTURE (oracle).  Stand-in so that the reference's
 * horizonator-lib.c compiles unmodified.  Only its texture branch (out of scope) uses
 * FreeImage; these entry points exist to satisfy the compiler and linker and fail if called. */
#pragma once
typedef unsigned char BYTE;
typedef struct FIBITMAP FIBITMAP;
typedef int FREE_IMAGE_FORMAT;
typedef int FREE_IMAGE_COLOR_TYPE;
#define FIF_UNKNOWN (-1)
#define FIC_PALETTE 3

FREE_IMAGE_FORMAT FreeImage_GetFileType(const char* filename, int size);
FIBITMAP* FreeImage_Load(FREE_IMAGE_FORMAT fif, const char* filename, int flags);
FREE_IMAGE_COLOR_TYPE FreeImage_GetColorType(FIBITMAP* dib);
FIBITMAP* FreeImage_ConvertTo24Bits(FIBITMAP* dib);
void FreeImage_Unload(FIBITMAP* dib);
unsigned FreeImage_GetWidth(FIBITMAP* dib);
unsigned FreeImage_GetHeight(FIBITMAP* dib);
unsigned FreeImage_GetBPP(FIBITMAP* dib);
unsigned FreeImage_GetPitch(FIBITMAP* dib);
BYTE* FreeImage_GetBits(FIBITMAP* dib);

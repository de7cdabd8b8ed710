// oracle/stub/FreeImage.h -- TEST INFRASTRUCTURE. Declarations only, so that
// raylib/loader/dll_loader.h:15-49 and raylib/render/image.cc:150-258 compile.
// FreeImage itself is not in this container; image file I/O is out of scope.
#pragma once
typedef unsigned char BYTE;
typedef int BOOL;
struct FIBITMAP;
enum FREE_IMAGE_FORMAT { FIF_UNKNOWN = -1, FIF_BMP = 0, FIF_JPEG = 2, FIF_PNG = 13, FIF_HDR = 26 };
extern "C" {
void FreeImage_Initialise(BOOL load_local_plugins_only);
void FreeImage_DeInitialise(void);
FREE_IMAGE_FORMAT FreeImage_GetFIFFromFilename(const char* filename);
FIBITMAP* FreeImage_Load(FREE_IMAGE_FORMAT fif, const char* filename, int flags);
FIBITMAP* FreeImage_ConvertToRGBAF(FIBITMAP* dib);
void FreeImage_Unload(FIBITMAP* dib);
BYTE* FreeImage_GetBits(FIBITMAP* dib);
unsigned FreeImage_GetWidth(FIBITMAP* dib);
unsigned FreeImage_GetHeight(FIBITMAP* dib);
unsigned FreeImage_GetPitch(FIBITMAP* dib);
FIBITMAP* FreeImage_ConvertTo32Bits(FIBITMAP* dib);
FIBITMAP* FreeImage_ConvertFromRawBits(BYTE* bits, int width, int height, int pitch, unsigned bpp,
	unsigned red_mask, unsigned green_mask, unsigned blue_mask, BOOL topdown);
BOOL FreeImage_Save(FREE_IMAGE_FORMAT fif, FIBITMAP* dib, const char* filename, int flags);
}

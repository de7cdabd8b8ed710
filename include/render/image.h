// Pixel (4 x f32 RGBA) and Image2D, the accumulation/output target of Raylib_Render
// and the storage behind textures and the sky panorama
// (reference: render/image.h:8-119, render/image.cc:12-134).
#pragma once

#include "raylib_types.h"
#include "core/int_types.h"
#include "core/vec3.h"
#include <vector>

struct Pixel
{
	float r, g, b, a;

	Pixel() : r(0.0f), g(0.0f), b(0.0f), a(1.0f) {}
	Pixel(float inR, float inG, float inB, float inA) : r(inR), g(inG), b(inB), a(inA) {}
	Pixel(float inR, float inG, float inB) : r(inR), g(inG), b(inB), a(1.0f) {}
	Pixel(uint8 inR, uint8 inG, uint8 inB, uint8 inA)
		: r((float)inR / 255.0f), g((float)inG / 255.0f), b((float)inB / 255.0f), a((float)inA / 255.0f) {}
	Pixel(uint32 argb)
		: r((float)((argb >> 16) & 0xff) / 255.0f)
		, g((float)((argb >> 8) & 0xff) / 255.0f)
		, b((float)(argb & 0xff) / 255.0f)
		, a((float)((argb >> 24) & 0xff) / 255.0f) {}

	uint32 ToUint32() const
	{
		const uint32 A = (uint32)(a * 255.0f) & 0xff, R = (uint32)(r * 255.0f) & 0xff;
		const uint32 G = (uint32)(g * 255.0f) & 0xff, B = (uint32)(b * 255.0f) & 0xff;
		return (A << 24) | (R << 16) | (G << 8) | B;
	}
	vec3 RGBToVec3() const { return vec3(r, g, b); }

	// Gamma 2.2 on all four channels, alpha included (reference quirk, image.h:79-83).
	Pixel LinearToSRGB() { const float k = 1.0f / 2.2f; return Pixel{ powf(r, k), powf(g, k), powf(b, k), powf(a, k) }; }
	Pixel SRGBToLinear() { const float k = 2.2f; return Pixel{ powf(r, k), powf(g, k), powf(b, k), powf(a, k) }; }
};

class Image2D
{
public:
	RAYLIB_API explicit Image2D();
	RAYLIB_API Image2D(uint32 width, uint32 height, const Pixel& fill);
	RAYLIB_API Image2D(uint32 width, uint32 height, uint32 argb = 0xff000000);

	// std::vector::resize semantics: existing pixels survive (reference: image.cc:27-32).
	RAYLIB_API void Reallocate(uint32 width, uint32 height, const Pixel& clearColor = Pixel(0xff000000));

	RAYLIB_API void SetPixel(int32 x, int32 y, const Pixel& value);
	RAYLIB_API void SetPixel(int32 x, int32 y, uint32 argb);

	RAYLIB_API void PostProcess();   // extended-Reinhard tone map + clamp + gamma, host-side

	uint32 GetWidth() const { return width; }
	uint32 GetHeight() const { return height; }
	Pixel GetPixel(int32 x, int32 y) const { return image[y * width + x]; }
	const std::vector<Pixel>& GetPixelArray() const { return image; }

	Image2D Clone() const;
	void DumpFloatRGBs(std::vector<float>& outArray) const;
	void DumpFloatRGBs(float* outArray) const;

	// B200 addition (non-virtual, no layout change): raw row-major storage for
	// bulk device<->host copies.
	Pixel* MutablePixels() { return image.data(); }

private:
	uint32 width;
	uint32 height;
	std::vector<Pixel> image;   // row-major, row 0 = top
};

namespace ImageIO
{
	RAYLIB_API Image2D* LoadImage2DFromFile(const char* filepath);
	bool WriteImage2DToDisk(Image2D* image, const char* filepath, EImageFileType fileType);
}

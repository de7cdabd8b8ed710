// Single-mip texture with the reference's sampling rule: nearest texel
// (int)((extent-1)*u), repeat wrap through fmodf, v flipped, optional gamma-2.2
// decode of all four channels (reference: render/texture.h:11-59, render/texture.cc:30-53).
// The device sampler in csrc/device implements the same rule on the flattened copy.
#pragma once

#include "raylib_types.h"
#include "render/image.h"
#include "core/int_types.h"

#include <vector>
#include <memory>

class Noncopyable
{
public:
	Noncopyable() = default;
	virtual ~Noncopyable() = default;
	Noncopyable(const Noncopyable&) = delete;
	Noncopyable& operator=(const Noncopyable&) = delete;
};

enum class ETextureFilter : uint8 { Nearest, Linear };
enum class ETextureWrap : uint8 { Clamp, Repeat };

struct SamplerState
{
	SamplerState() : filter(ETextureFilter::Linear), wrap(ETextureWrap::Repeat), bSRGB(false) {}
	ETextureFilter filter;   // recorded, but sampling is always nearest (as in the reference)
	ETextureWrap wrap;       // recorded, but sampling always repeats
	bool bSRGB;
};

class Texture2D : public Noncopyable
{
public:
	RAYLIB_API static Texture2D* CreateFromImage2D(std::shared_ptr<Image2D> inImage);
	static Texture2D* CreateSolidColor(const Pixel& inColor);

	Texture2D(uint32 numMipmaps);

	void SetData(uint32 mipLevel, std::shared_ptr<Image2D> image);
	void SetSamplerState(const SamplerState& inSampler) { sampler = inSampler; }

	RAYLIB_API Pixel Sample(float u, float v);

private:
	friend struct RtSceneFlattener;
	std::vector<std::shared_ptr<Image2D>> mipmaps;
	SamplerState sampler;
};

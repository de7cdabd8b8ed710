// Single-mip texture with the reference's sampling rule: nearest texel
// (int)((extent-1)*u), repeat wrap through fmodf, v flipped, optional gamma-2.2
// decode of all four channels (reference: render/texture.h:11-59, render/texture.cc:30-53).
// The device sampler in csrc/device implements the same rule on the flattened copy.
#pragma once
#include <memory>
#include <vector>
#include "core/int_types.h"
#include "render/image.h"
#include "raylib_types.h"

// base of the classes that must not be copied (reference: core/noncopyable.h; that header forwards here)
class Noncopyable
{
public:
	Noncopyable() = default;
	virtual ~Noncopyable() = default;
	Noncopyable(const Noncopyable&) = delete;
	Noncopyable& operator=(const Noncopyable&) = delete;
};

enum class ETextureWrap : uint8 { Clamp, Repeat };
enum class ETextureFilter : uint8 { Nearest, Linear };

// filter and wrap are recorded only: sampling is always nearest + repeat, as in the reference; bSRGB switches the
// pow(x, 2.2) decode of all four channels
struct SamplerState
{
	ETextureFilter filter = ETextureFilter::Linear;
	ETextureWrap   wrap = ETextureWrap::Repeat;
	bool           bSRGB = false;
};

class Texture2D : public Noncopyable
{
	friend struct RtSceneFlattener;
	std::vector<std::shared_ptr<Image2D>> mipmaps;       // level 0 is the one that is sampled
	SamplerState sampler;

public:
	Texture2D(uint32 numMipmaps);
	void SetSamplerState(const SamplerState& inSampler) { sampler = inSampler; }
	void SetData(uint32 mipLevel, std::shared_ptr<Image2D> image);
	RAYLIB_API Pixel Sample(float u, float v);

	// one-level textures: from an image (shared with the caller), or 1x1 of a colour
	static Texture2D* CreateSolidColor(const Pixel& inColor);
	RAYLIB_API static Texture2D* CreateFromImage2D(std::shared_ptr<Image2D> inImage);
};

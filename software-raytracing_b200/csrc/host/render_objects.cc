// render_objects.cc -- host side of Image2D, Texture2D and the material classes.
// Images are plain host storage (upload/download targets); materials are parameter holders whose
// scattering code lives in csrc/device/rt_shade.cuh.  Reference behaviour: raylib/render/image.cc:12-134,
// raylib/render/texture.cc:4-53, raylib/render/material.cc:342-415.
#include "gpu_state.h"
#include "render/image.h"
#include "render/texture.h"
#include "render/material.h"
#include "render/renderer.h"
#include "core/logger.h"
#include "host_internal.h"

#include <cstdio>
#include <cstring>

// ---------------------------------------------------------------------------------------------
// Image2D

Image2D::Image2D() { Reallocate(0, 0); }
Image2D::Image2D(uint32 inWidth, uint32 inHeight, const Pixel& fill) { Reallocate(inWidth, inHeight, fill); }
Image2D::Image2D(uint32 inWidth, uint32 inHeight, uint32 argb) : Image2D(inWidth, inHeight, Pixel(argb)) {}

void Image2D::Reallocate(uint32 inWidth, uint32 inHeight, const Pixel& clearColor)
{
	RtGpu::ForgetHostImage(this);      // the storage may move: drop its page-lock (taken again by the next render into it)
	width = inWidth;
	height = inHeight;
	image.resize((size_t)width * height, clearColor);
}

void Image2D::SetPixel(int32 x, int32 y, const Pixel& value) { image[(size_t)y * width + x] = value; }
void Image2D::SetPixel(int32 x, int32 y, uint32 argb) { image[(size_t)y * width + x] = Pixel(argb); }

// Extended-Reinhard luminance tone map, clamp to white, gamma 2.2 (image.cc:44-103).  Host-side
// post step on the finished frame; not part of the per-ray path.
void Image2D::PostProcess()
{
	const size_t count = (size_t)width * height;
	const vec3 lumaWeights(0.2126f, 0.7152f, 0.0722f);

	float maxWhite = 1.0f;
	for (size_t i = 0; i < count; ++i)
	{
		const float luma = dot(vec3(image[i].r, image[i].g, image[i].b), lumaWeights);
		if (maxWhite < luma) maxWhite = luma;
	}
	LOG("Max white luminance: %f", maxWhite);

	for (size_t i = 0; i < count; ++i)
	{
		vec3 rgb(image[i].r, image[i].g, image[i].b);
		const float lumaOld = dot(rgb, lumaWeights);
		if (lumaOld <= 0.0001f) rgb = vec3(0.0f);
		else
		{
			const float numerator = lumaOld * (1.0f + (lumaOld / (maxWhite * maxWhite)));
			const float lumaNew = numerator / (1.0f + lumaOld);
			rgb = rgb * (lumaNew / lumaOld);
		}
		rgb = min(vec3(1.0f), rgb);
		rgb = pow(rgb, 1.0f / 2.2f);
		image[i].r = rgb.x; image[i].g = rgb.y; image[i].b = rgb.z;
	}
}

Image2D Image2D::Clone() const
{
	Image2D copy;
	copy.width = width;
	copy.height = height;
	copy.image = image;
	return copy;
}

void Image2D::DumpFloatRGBs(std::vector<float>& outArray) const
{
	outArray.assign((size_t)3 * width * height, 0.0f);
	DumpFloatRGBs(outArray.data());
}

void Image2D::DumpFloatRGBs(float* outArray) const
{
	const size_t count = (size_t)width * height;
	for (size_t i = 0; i < count; ++i)
	{
		outArray[3 * i + 0] = image[i].r;
		outArray[3 * i + 1] = image[i].g;
		outArray[3 * i + 2] = image[i].b;
	}
}

// ImageIO (file codecs) lives in image_codecs.cc

// ---------------------------------------------------------------------------------------------
// Texture2D

Texture2D* Texture2D::CreateFromImage2D(std::shared_ptr<Image2D> inImage)
{
	Texture2D* texture = new Texture2D(1);
	texture->SetData(0, inImage);
	return texture;
}

Texture2D* Texture2D::CreateSolidColor(const Pixel& inColor)
{
	std::shared_ptr<Image2D> image = std::make_shared<Image2D>();
	image->Reallocate(1, 1, inColor);
	return CreateFromImage2D(image);
}

Texture2D::Texture2D(uint32 numMipmaps) { mipmaps.resize(numMipmaps); }

void Texture2D::SetData(uint32 mipLevel, std::shared_ptr<Image2D> image) { mipmaps[mipLevel] = image; }

// Host-side texel lookup with the same addressing rule the device sampler uses.  It is a data
// accessor for clients (e.g. GetAlbedo); rendering samples the uploaded copy on the GPU.
Pixel Texture2D::Sample(float u, float v)
{
	if (mipmaps.empty() || !mipmaps[0]) return Pixel(0.0f, 0.0f, 0.0f, 0.0f);
	u = fmodf(u, 1.0f); if (u < 0.0f) u += 1.0f;
	v = fmodf(v, 1.0f); if (v < 0.0f) v += 1.0f; v = 1.0f - v;
	if (std::isnan(u) || std::isinf(u)) u = 0.0f;
	if (std::isnan(v) || std::isinf(v)) v = 0.0f;
	const Image2D& mip = *mipmaps[0];
	const int32 x = (int32)((mip.GetWidth() - 1) * u);
	const int32 y = (int32)((mip.GetHeight() - 1) * v);
	Pixel px = mip.GetPixel(x, y);
	if (sampler.bSRGB) px = px.SRGBToLinear();
	return px;
}

// ---------------------------------------------------------------------------------------------
// Materials: Scatter / ScatteringPdf are device-only.

bool Lambertian::Scatter(const ray&, const HitResult&, vec3&, ray&, float&) const { return RtHostQueryUnsupported("Lambertian::Scatter"); }
float Lambertian::ScatteringPdf(const HitResult& hitResult, const vec3&, const vec3& Wi) const
{
	return std::max(0.0f, dot(hitResult.n, Wi)) / BRDF::PI;
}
bool Metal::Scatter(const ray&, const HitResult&, vec3&, ray&, float&) const { return RtHostQueryUnsupported("Metal::Scatter"); }
bool Dielectric::Scatter(const ray&, const HitResult&, vec3&, ray&, float&) const { return RtHostQueryUnsupported("Dielectric::Scatter"); }
bool Mirror::Scatter(const ray&, const HitResult&, vec3&, ray&, float&) const { return RtHostQueryUnsupported("Mirror::Scatter"); }
bool MicrofacetMaterial::Scatter(const ray&, const HitResult&, vec3&, ray&, float&) const { return RtHostQueryUnsupported("MicrofacetMaterial::Scatter"); }
float MicrofacetMaterial::ScatteringPdf(const HitResult&, const vec3&, const vec3&) const
{
	RtHostQueryUnsupported("MicrofacetMaterial::ScatteringPdf");
	return 0.0f;
}

vec3 MicrofacetMaterial::Emitted(const HitResult& hitResult, const vec3&) const
{
	if (emissiveTexture)
	{
		// reference quirk kept (material.cc:345-346): lookup at (u,u), blue channel broadcast
		const Pixel px = emissiveTexture->Sample(hitResult.paramU, hitResult.paramU);
		return vec3(px.b);
	}
	return emissiveFallback;
}

bool MicrofacetMaterial::IsMirrorLike(float paramU, float paramV) const
{
	const float roughness = roughnessTexture ? roughnessTexture->Sample(paramU, paramV).r : roughnessFallback;
	return roughness < 0.1f;
}

vec3 MicrofacetMaterial::GetAlbedo(float paramU, float paramV) const
{
	if (!albedoTexture) return albedoFallback;
	const Pixel px = albedoTexture->Sample(paramU, paramV);
	return px.RGBToVec3() * px.a;
}

bool MicrofacetMaterial::AlphaTest(float texcoordU, float texcoordV) const
{
	if (!albedoTexture) return true;
	return albedoTexture->Sample(texcoordU, texcoordV).a >= 0.5f;
}

vec3 MicrofacetMaterial::GetMicrosurfaceNormal(const HitResult& hitResult) const
{
	if (!normalmapTexture) return vec3(0.0f, 0.0f, 1.0f);
	const vec3 encoded = normalmapTexture->Sample(hitResult.paramU, hitResult.paramV).RGBToVec3();
	return normalize(2.0f * encoded - 1.0f);
}

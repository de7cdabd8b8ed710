// Material classes clients attach to primitives
// (reference: render/material.h:16-270, render/material.cc:195-431).
// Scattering is evaluated by the device shade kernels, which read a flattened
// parameter record per material; the host classes hold parameters and textures.
// Virtual-function order and field order are ABI.
#pragma once

#include "raylib_types.h"
#include "core/random.h"
#include "render/brdf.h"
#include "render/image.h"
#include "render/texture.h"
#include "geom/hit.h"

#include <algorithm>
#include <memory>

class Material
{
public:
	// Device-only in this library: the host entry point reports an error and returns false.
	RAYLIB_API virtual bool Scatter(
		const ray& inPathRay, const HitResult& inHitResult,
		vec3& outReflectance, ray& outScatteredRay, float& outPdf) const = 0;

	RAYLIB_API virtual vec3 Emitted(const HitResult& hitResult, const vec3& Wo) const { return vec3(0.0f); }

	RAYLIB_API virtual float ScatteringPdf(const HitResult& hitResult, const vec3& Wo, const vec3& Wi) const
	{
		return 1.0f / BRDF::PI;
	}

	virtual bool IsMirrorLike(float paramU, float paramV) const { return false; }
	RAYLIB_API virtual vec3 GetAlbedo(float paramU, float paramV) const { return vec3(0.0f); }
	RAYLIB_API virtual bool AlphaTest(float paramU, float paramV) const { return true; }
	RAYLIB_API virtual vec3 GetMicrosurfaceNormal(const HitResult& hitResult) const { return vec3(0.0f, 0.0f, 1.0f); }
};

class DiffuseLight : public Material
{
public:
	RAYLIB_API DiffuseLight(const vec3& inIntensity) : intensity(inIntensity) {}

	RAYLIB_API bool Scatter(const ray&, const HitResult&, vec3&, ray&, float&) const override { return false; }
	RAYLIB_API vec3 Emitted(const HitResult& hitResult, const vec3& Wo) const override { return intensity; }

	vec3 intensity;
};

class Lambertian : public Material
{
public:
	RAYLIB_API Lambertian(const vec3& inAlbedo) { albedo = saturate(inAlbedo); }

	RAYLIB_API bool Scatter(const ray& inRay, const HitResult& inResult,
		vec3& outReflectance, ray& outScatteredRay, float& outPdf) const override;
	RAYLIB_API float ScatteringPdf(const HitResult& hitResult, const vec3& Wo, const vec3& Wi) const override;
	RAYLIB_API virtual vec3 GetAlbedo(float paramU, float paramV) const override { return albedo; }

private:
	friend struct RtSceneFlattener;
	vec3 albedo;
};

class Metal : public Material
{
public:
	RAYLIB_API Metal(const vec3& inAlbedo, float inFuzziness = 0.0f)
		: albedo(inAlbedo), fuzziness(std::max(0.0f, std::min(1.0f, inFuzziness))) {}

	RAYLIB_API bool Scatter(const ray& inRay, const HitResult& inResult,
		vec3& outReflectance, ray& outScatteredRay, float& outPdf) const override;
	RAYLIB_API virtual vec3 GetAlbedo(float paramU, float paramV) const { return albedo; }

	vec3 albedo;
	float fuzziness;
};

class Dielectric : public Material
{
public:
	RAYLIB_API Dielectric(float indexOfRefraction, const vec3& inTransmissionFilter = vec3(1.0f))
		: ref_idx(indexOfRefraction), transmissionFilter(inTransmissionFilter) {}

	RAYLIB_API bool Scatter(const ray& inRay, const HitResult& inResult,
		vec3& outReflectance, ray& outScatteredRay, float& outPdf) const override;
	virtual bool IsMirrorLike(float paramU, float paramV) const override { return true; }

	float ref_idx;
	vec3 transmissionFilter;
};

class Mirror : public Material
{
public:
	Mirror(const vec3& inBaseColor = vec3(1.0f)) : baseColor(inBaseColor) {}

	virtual bool Scatter(const ray& inRay, const HitResult& inResult,
		vec3& outReflectance, ray& outScatteredRay, float& outPdf) const override;
	virtual float ScatteringPdf(const HitResult& hitResult, const vec3& Wo, const vec3& Wi) const { return 1.0f; }
	virtual bool IsMirrorLike(float paramU, float paramV) const override { return true; }

	vec3 baseColor;
};

// Beckmann microfacet material with optional albedo / normal / roughness /
// metallic / emissive textures and a 0.5 alpha cut-out on the albedo texture.
class MicrofacetMaterial : public Material
{
public:
	static MicrofacetMaterial* FromConstants(const vec3& inAlbedo, const float inRoughness,
		const float inMetallic, const vec3& inEmissive)
	{
		MicrofacetMaterial* m = new MicrofacetMaterial;
		m->albedoFallback = inAlbedo;
		m->roughnessFallback = inRoughness;
		m->metallicFallback = inMetallic;
		m->emissiveFallback = inEmissive;
		return m;
	}

	RAYLIB_API MicrofacetMaterial()
		: albedoTexture(nullptr), normalmapTexture(nullptr), roughnessTexture(nullptr)
		, metallicTexture(nullptr), emissiveTexture(nullptr)
		, albedoFallback(vec3(0.5f)), roughnessFallback(1.0f), metallicFallback(0.0f), emissiveFallback(vec3(0.0f)) {}

	RAYLIB_API void SetAlbedoTexture(std::shared_ptr<Image2D> inImage)
	{
		ReplaceTexture(albedoTexture, inImage);
		SamplerState srgb;
		srgb.bSRGB = true;
		albedoTexture->SetSamplerState(srgb);
	}
	RAYLIB_API void SetNormalTexture(std::shared_ptr<Image2D> inImage) { ReplaceTexture(normalmapTexture, inImage); }
	RAYLIB_API void SetRoughnessTexture(std::shared_ptr<Image2D> inImage) { ReplaceTexture(roughnessTexture, inImage); }
	RAYLIB_API void SetMetallicTexture(std::shared_ptr<Image2D> inImage) { ReplaceTexture(metallicTexture, inImage); }
	RAYLIB_API void SetEmissiveTexture(std::shared_ptr<Image2D> inImage) { ReplaceTexture(emissiveTexture, inImage); }

	void SetAlbedoFallback(const vec3& inAlbedo) { albedoFallback = saturate(inAlbedo); }
	void SetRoughnessFallback(float inRoughness) { roughnessFallback = std::min(1.0f, std::max(0.0f, inRoughness)); }
	void SetMetallicFallback(float inMetallic) { metallicFallback = std::min(1.0f, std::max(0.0f, inMetallic)); }
	void SetEmissiveFallback(const vec3& inEmissive) { emissiveFallback = inEmissive; }

	RAYLIB_API bool Scatter(const ray& inRay, const HitResult& inResult,
		vec3& outReflectance, ray& outScatteredRay, float& outPdf) const override;
	RAYLIB_API vec3 Emitted(const HitResult& hitResult, const vec3& Wo) const override;
	RAYLIB_API float ScatteringPdf(const HitResult& hitResult, const vec3& Wo, const vec3& Wi) const override;
	virtual bool IsMirrorLike(float paramU, float paramV) const override;
	RAYLIB_API virtual vec3 GetAlbedo(float paramU, float paramV) const override;
	RAYLIB_API virtual bool AlphaTest(float texcoordU, float texcoordV) const override;
	RAYLIB_API virtual vec3 GetMicrosurfaceNormal(const HitResult& hitResult) const override;

private:
	friend struct RtSceneFlattener;
	static void ReplaceTexture(Texture2D*& slot, std::shared_ptr<Image2D> image)
	{
		if (slot) delete slot;
		slot = Texture2D::CreateFromImage2D(image);
	}

	Texture2D* albedoTexture;
	Texture2D* normalmapTexture;
	Texture2D* roughnessTexture;
	Texture2D* metallicTexture;
	Texture2D* emissiveTexture;

	vec3 albedoFallback;
	float roughnessFallback;
	float metallicFallback;
	vec3 emissiveFallback;
};

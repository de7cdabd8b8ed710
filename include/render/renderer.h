// Renderer facade (reference: render/renderer.h:12-29).  RenderScene uploads the
// flattened scene if needed, runs the CUDA wavefront path tracer and copies the
// frame back into the Image2D.  DenoiseScene needs OpenImageDenoise, which is a
// Windows-only prebuilt dependency of the reference; it reports "unsupported".
#pragma once
#include "core/vec3.h"
#include "core/int_types.h"
#include "raylib_types.h"

class Image2D; class Scene; class Camera; class Hitable;

class Renderer
{
public:
	// blocking; outImage is resized to the settings' viewport when it differs (render/renderer.cc:292-296)
	void RenderScene(const RendererSettings* settings, const Scene* world, const Camera* camera, Image2D* outImage);

	// false off Windows, like the reference (render/renderer.cc:28-33): DenoiseScene then returns false and leaves the output alone
	static bool IsDenoiserSupported();
	bool DenoiseScene(Image2D* mainImage, bool bMainImageHDR, Image2D* albedoImage, Image2D* normalImage, Image2D* outDenoisedImage);
};

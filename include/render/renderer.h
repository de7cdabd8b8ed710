// Renderer facade (reference: render/renderer.h:12-29).  RenderScene uploads the
// flattened scene if needed, runs the CUDA wavefront path tracer and copies the
// frame back into the Image2D.  DenoiseScene needs OpenImageDenoise, which is a
// Windows-only prebuilt dependency of the reference; it reports "unsupported".
#pragma once

#include "raylib_types.h"
#include "core/int_types.h"
#include "core/vec3.h"

class Hitable;
class Camera;
class Scene;
class Image2D;

class Renderer
{
public:
	static bool IsDenoiserSupported();

	void RenderScene(const RendererSettings* settings, const Scene* world, const Camera* camera, Image2D* outImage);

	bool DenoiseScene(Image2D* mainImage, bool bMainImageHDR,
		Image2D* albedoImage, Image2D* normalImage, Image2D* outDenoisedImage);
};

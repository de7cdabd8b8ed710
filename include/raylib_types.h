// raylib_types.h -- handle, enum and settings types of the raylib C ABI.
// Layout-identical to the reference (raylib/raylib_types.h:13-57): handles are
// pointer-sized integers, RendererSettings is a 24-byte POD whose C# mirror is
// gui-app/gui-app/RaylibWrapper.cs:27-38.
#pragma once

#include <stdint.h>

#if defined(_MSC_VER)
  #ifdef RAYLIB_EXPORTS
    #define RAYLIB_API __declspec(dllexport)
  #else
    #define RAYLIB_API __declspec(dllimport)
  #endif
#else
  #define RAYLIB_API __attribute__((visibility("default")))
#endif

typedef uintptr_t OBJModelHandle;
typedef uintptr_t ImageHandle;
typedef uintptr_t SceneHandle;
typedef uintptr_t SceneElementHandle;   // = Hitable*
typedef uintptr_t CameraHandle;

// What Raylib_Render writes into the image.  0 runs the path tracer, the rest
// are one-ray-per-pixel debug views (reference: render/renderer.cc:62-111).
enum ERenderMode
{
	RAYLIB_RENDERMODE_Default            = 0,
	RAYLIB_RENDERMODE_Albedo             = 1,
	RAYLIB_RENDERMODE_SurfaceNormal      = 2,   // world space; README calls it "VertexNormal"
	RAYLIB_RENDERMODE_MicrosurfaceNormal = 3,
	RAYLIB_RENDERMODE_Texcoord           = 4,
	RAYLIB_RENDERMODE_Emission           = 5,
	RAYLIB_RENDERMODE_Reflectance        = 6,

	RAYLIB_RENDERMODE_MAX
};

enum EImageFileType
{
	RAYLIB_IMAGEFILETYPE_Bitmap = 0,
	RAYLIB_IMAGEFILETYPE_Jpg    = 1,
	RAYLIB_IMAGEFILETYPE_Png    = 2,

	RAYLIB_IMAGEFILETYPE_MAX
};

struct RendererSettings
{
	uint32_t viewportWidth;
	uint32_t viewportHeight;
	int32_t  samplesPerPixel;    // values < 1 render one sample
	int32_t  maxPathLength;      // paths reaching this depth contribute zero
	float    rayTMin;            // lower bound of every ray query, camera rays included
	uint32_t renderMode = RAYLIB_RENDERMODE_Default;

	inline float getViewportAspectWH() const { return (float)viewportWidth / (float)viewportHeight; }
};

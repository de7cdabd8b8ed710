// raylib.h -- the C ABI of the B200 raylib drop-in.
//
// Same 33 entry points, argument meaning and return conventions as the
// reference's raylib/raylib.h:23-149 (implemented there in raylib/raylib.cc).
// Everything the reference's FFI users bind (gui-app/gui-app/RaylibWrapper.cs:43-145)
// resolves against this library unchanged.  What differs is behind
// Raylib_FinalizeScene / Raylib_Render: the scene is flattened to GPU records and
// the frame is produced by the sm_100a wavefront path tracer.
// There is no CPU rendering path: Raylib_Render fails loudly without a CUDA device.
// Raylib_Render spreads a frame over every visible GPU of the box (raylib_b200.h: RaylibB200_SetDevices), like the
// reference spreads it over every core.  Scene elements may be Sphere, Cube, Triangle, StaticMesh, BVHNode and raw
// HitableList objects holding those primitives; user-defined Hitable / Material subclasses are refused at first render.
#pragma once

#include "raylib_types.h"
#include "core/stat.h"
#include "core/logger.h"
#include <stdint.h>

extern "C" {

// ---- library lifetime (raylib.cc:25-51). Both return 1 on success... Terminate returns 0 in the reference; kept.
RAYLIB_API int32_t Raylib_Initialize();
RAYLIB_API int32_t Raylib_Terminate();

// ---- media (raylib.cc:56-113). Host-side.  The reference parses OBJ with tinyobjloader and images with FreeImage.dll;
// this build carries its own Wavefront OBJ/MTL importer (csrc/host/obj_loader.cc: the reference's conversion rules, faces
// kept as arrays and flattened without Triangle/BVHNode objects) and PNG / BMP / TGA / JPEG / Radiance HDR / PNM decoders
// (csrc/host/image_codecs.cc, jpeg_codec.cc: baseline and progressive Huffman JPEG, libjpeg's bytes).  A file that cannot
// be read or decoded returns NULL, as in the reference (stderr names the file).
RAYLIB_API OBJModelHandle Raylib_LoadOBJModel(const char* objPath);
RAYLIB_API void Raylib_TransformOBJModel(OBJModelHandle objModel,
	float translationX, float translationY, float translationZ,
	float yaw, float pitch, float roll,
	float scaleX, float scaleY, float scaleZ);
RAYLIB_API void Raylib_FinalizeOBJModel(OBJModelHandle objModel);
RAYLIB_API int32_t Raylib_UnloadOBJModel(OBJModelHandle objHandle);
RAYLIB_API ImageHandle Raylib_LoadImage(const char* filepath);

// ---- scene (raylib.cc:205-283). Elements are caller-owned Hitable* objects.
RAYLIB_API SceneHandle Raylib_CreateScene();
RAYLIB_API void Raylib_AddSceneElement(SceneHandle scene, SceneElementHandle element);
RAYLIB_API void Raylib_AddOBJModelToScene(SceneHandle scene, OBJModelHandle objModel);
RAYLIB_API void Raylib_SetSkyPanorama(SceneHandle scene, ImageHandle skyImage);
RAYLIB_API void Raylib_SetSunIlluminance(SceneHandle scene, float r, float g, float b);
RAYLIB_API void Raylib_SetSunDirection(SceneHandle scene, float x, float y, float z);
RAYLIB_API void Raylib_FinalizeScene(SceneHandle scene);      // builds the BVH; scene is immutable afterwards
RAYLIB_API int32_t Raylib_DestroyScene(SceneHandle sceneHandle);

// ---- camera (raylib.cc:118-179)
RAYLIB_API CameraHandle Raylib_CreateCamera();
RAYLIB_API void Raylib_CameraSetPosition(CameraHandle camera, float x, float y, float z);
RAYLIB_API void Raylib_CameraSetLookAt(CameraHandle camera, float tx, float ty, float tz);
RAYLIB_API void Raylib_CameraSetPerspective(CameraHandle camera, float fovY_degrees, float aspectWH);
RAYLIB_API void Raylib_CameraSetLens(CameraHandle camera, float aperture, float focalDistance);
RAYLIB_API void Raylib_CameraSetMotion(CameraHandle camera, float beginTime, float endTime);
RAYLIB_API void Raylib_CameraCopy(CameraHandle srcCamera, CameraHandle dstCamera);
RAYLIB_API int32_t Raylib_DestroyCamera(CameraHandle cameraHandle);

// ---- images (raylib.cc:181-203). DumpImageData: caller provides 3*W*H floats, row-major RGB, row 0 on top.
RAYLIB_API ImageHandle Raylib_CreateImage(uint32_t width, uint32_t height);
RAYLIB_API void Raylib_DumpImageData(ImageHandle image, float* outDest);
RAYLIB_API int32_t Raylib_DestroyImage(ImageHandle imageHandle);

// ---- rendering (raylib.cc:231-293). Raylib_Render is synchronous; it resizes
// outMainImage to the settings viewport when they differ (renderer.cc:292-296).
RAYLIB_API void Raylib_Render(const RendererSettings* settings,
	SceneHandle scene, CameraHandle camera, ImageHandle outMainImage);
RAYLIB_API int32_t Raylib_Denoise(ImageHandle inMainImage, int32_t bMainImageHDR,
	ImageHandle inAlbedoImage, ImageHandle inNormalImage, ImageHandle outDenoisedImage);
RAYLIB_API void Raylib_PostProcess(ImageHandle image);
RAYLIB_API int32_t Raylib_IsDenoiserSupported();

// ---- utilities (raylib.cc:298-331)
RAYLIB_API const char* Raylib_GetRenderModeString(uint32_t auxMode);
RAYLIB_API int32_t Raylib_WriteImageToDisk(ImageHandle image, const char* filepath, uint32_t fileType);
RAYLIB_API void Raylib_FlushLogThread();

} // extern "C"

// abi_layout_check.cc -- pins the object sizes of the reference's public classes (measured with
// g++ on the reference headers, SURVEY.md section 7 "ABI breadth").  A client binary compiled
// against either header set must agree with this library on every one of them.
#include "geom/scene.h"
#include "geom/primitives.h"
#include "render/camera.h"
#include "render/material.h"
#include "render/image.h"
#include "raylib_types.h"

static_assert(sizeof(vec3) == 12, "vec3");
static_assert(sizeof(ray) == 28, "ray");
static_assert(sizeof(AABB) == 24, "AABB");
static_assert(sizeof(HitResult) == 72, "HitResult");
static_assert(sizeof(BVHNode) == 48, "BVHNode");
static_assert(sizeof(Sphere) == 32, "Sphere");
static_assert(sizeof(Cube) == 56, "Cube");
static_assert(sizeof(Triangle) == 152, "Triangle");
static_assert(sizeof(StaticMesh) == 80, "StaticMesh");
static_assert(sizeof(Scene) == 80, "Scene");
static_assert(sizeof(Camera) == 128, "Camera");
static_assert(sizeof(Lambertian) == 24, "Lambertian");
static_assert(sizeof(Metal) == 24, "Metal");
static_assert(sizeof(Dielectric) == 24, "Dielectric");
static_assert(sizeof(Mirror) == 24, "Mirror");
static_assert(sizeof(DiffuseLight) == 24, "DiffuseLight");
static_assert(sizeof(MicrofacetMaterial) == 80, "MicrofacetMaterial");
static_assert(sizeof(RendererSettings) == 24, "RendererSettings");
static_assert(sizeof(Pixel) == 16, "Pixel");
static_assert(sizeof(Image2D) == 32, "Image2D");

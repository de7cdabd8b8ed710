// Product build of scenes.cc: the library seeds its own BVH split-axis stream, nothing to do here.
extern "C" __attribute__((visibility("default"))) void scene_hook_before_bvh_build(void) {}
extern "C" __attribute__((visibility("default"))) void scene_hook_on_destroy(unsigned long) {}

// Scene ingestion into the flat format: the step immediately before the rendering path.
//   import_mitsuba   src/common/importer/mitsuba.rs (schema, get_camera :685-710, shape generators :20-150)
//                    + src/pathtracer/importer/mitsuba.rs (materials :84-181, shapes :183-330, emitters :365-418)
//   import_gltf      src/common/importer/gltf.rs + src/pathtracer/importer/gltf.rs
//   import_scene     src/common/importer/mod.rs:6-25 (dispatch on the file extension)
// Each fills a SceneBuilder exactly as the reference fills its RenderScene (same mesh / triangle / light
// order, so the SAH build and the light indices come out the same) and returns the camera.
#pragma once
#include <string>

#include "scene_builder.hpp"

namespace ptrs_host {

struct ImportOptions {
  int res_w = 640, res_h = 480;   // -r WxH; DEFAULT_RESOLUTION is 640x480 (src/common/mod.rs:14)
  bool default_lights = false;    // --default_lights (glTF only)
  // `<emitter type="sunsky"/>` maps to CARGO_MANIFEST_DIR/data/abandoned_tank_farm_04_1k.hdr in the reference
  // (importer/mitsuba.rs:400-418).  That file does not ship with this repo: give its path here; when empty the
  // seeded synthetic sky of procedural.hpp is used instead.
  std::string sunsky_hdr;
};

PtrsCamera import_mitsuba(const std::string& path, const ImportOptions& opt, SceneBuilder& b);
PtrsCamera import_gltf(const std::string& path, const ImportOptions& opt, SceneBuilder& b);
PtrsCamera import_scene(const std::string& path, const ImportOptions& opt, SceneBuilder& b);

// heck::SnakeCase as the importer applies it to parameter names ("intIOR" -> "int_ior")
std::string snake_case(const std::string& s);
// math.rs:141-147
float inverse_gamma_correct(float v);

}  // namespace ptrs_host

// Counterpart of the reference's examples/small_scene.rs: one headless Cornell-box render through the
// C++ mirror of the reference interface, written as a binary PPM.
//   make -C examples && ./examples/small_scene out.ppm [width height spp max_depth]
#include <cstdio>
#include <cstdlib>

#include "../pathtracer_rs_b200/host/pathtracer.hpp"
#include "../pathtracer_rs_b200/host/procedural.hpp"

int main(int argc, char** argv) {
  const char* out = argc > 1 ? argv[1] : "render.ppm";
  const int w = argc > 3 ? std::atoi(argv[2]) : 512, h = argc > 3 ? std::atoi(argv[3]) : 512;
  const int spp = argc > 4 ? std::atoi(argv[4]) : 16, depth = argc > 5 ? std::atoi(argv[5]) : 15;
  try {
    ptrs_host::SceneBuilder b;
    ptrs_host::build_cornell(b, nullptr, 0, 0);
    ptrs_host::FlatScene flat = b.finalize(4);
    ptrs::Camera camera(ptrs_host::cornell_camera(w, h));
    ptrs::RenderScene scene(flat);
    const float radius[2] = {2.f, 2.f};
    ptrs::SamplerBuilder sampler(spp, camera.film.get_sample_bounds(radius));
    ptrs::PathIntegrator integrator(sampler, depth, false);
    integrator.preprocess(scene);
    integrator.render(camera, scene);
    PtrsStats st = integrator.stats(scene);
    std::printf("rendering took: %.1f ms (%llu camera paths, %.1f M samples/s)\n", st.ms_total, (unsigned long long)st.camera_paths,
                st.camera_paths / st.ms_total / 1e3);
    std::vector<uint8_t> img = camera.film.to_rgba_image();
    FILE* f = std::fopen(out, "wb");
    std::fprintf(f, "P6\n%d %d\n255\n", w, h);
    for (size_t i = 0; i < (size_t)w * h; ++i) std::fwrite(&img[4 * i], 1, 3, f);
    std::fclose(f);
  } catch (const ptrs::Error& e) {
    std::fprintf(stderr, "ptrs error %d: %s\n", e.code, e.what());
    return 1;
  }
  return 0;
}

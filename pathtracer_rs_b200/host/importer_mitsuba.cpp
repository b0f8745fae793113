// Mitsuba 0.5 XML -> SceneBuilder.  Follows src/common/importer/mitsuba.rs (what is parsed) and
// src/pathtracer/importer/mitsuba.rs (what it becomes); unsupported constructs throw where the reference panics.
#include <array>
#include <cmath>
#include <cstdlib>
#include <map>
#include <sstream>
#include <stdexcept>

#include "image_io.hpp"
#include "importers.hpp"
#include "procedural.hpp"
#include "xml_lite.hpp"

namespace ptrs_host {

std::string snake_case(const std::string& s) {
  std::string out;
  auto is_up = [](char c) { return c >= 'A' && c <= 'Z'; };
  auto is_lo = [](char c) { return (c >= 'a' && c <= 'z') || (c >= '0' && c <= '9'); };
  for (size_t i = 0; i < s.size(); ++i) {
    const char c = s[i];
    if (!is_up(c) && !is_lo(c)) {  // separator
      if (!out.empty() && out.back() != '_') out.push_back('_');
      continue;
    }
    if (is_up(c) && i > 0) {
      const char p = s[i - 1];
      const bool next_lower = i + 1 < s.size() && s[i + 1] >= 'a' && s[i + 1] <= 'z';
      if ((is_lo(p) || (is_up(p) && next_lower)) && !out.empty() && out.back() != '_') out.push_back('_');
    }
    out.push_back(is_up(c) ? (char)(c - 'A' + 'a') : c);
  }
  while (!out.empty() && out.back() == '_') out.pop_back();
  return out;
}

float inverse_gamma_correct(float v) {
  if (v <= 0.04045f) return v * 1.0f / 12.92f;
  return std::pow((v + 0.055f) * 1.0f / 1.055f, 2.4f);
}

namespace {

[[noreturn]] void bad(const std::string& what) { throw std::runtime_error("mitsuba import: " + what); }

float parse_float(const std::string& s) {
  char* end = nullptr;
  const float v = std::strtof(s.c_str(), &end);
  if (end == s.c_str()) bad("cannot parse number '" + s + "'");
  return v;
}
std::vector<float> parse_floats(const std::string& s) {  // separators: spaces and / or commas
  std::vector<float> out;
  std::string tok;
  for (size_t i = 0; i <= s.size(); ++i) {
    const char c = i < s.size() ? s[i] : ' ';
    if (c == ' ' || c == ',' || c == '\t' || c == '\n') {
      if (!tok.empty()) out.push_back(parse_float(tok));
      tok.clear();
    } else {
      tok.push_back(c);
    }
  }
  return out;
}
const std::string& need_attr(const XmlNode& n, const char* key) {
  const std::string* v = n.attr(key);
  if (!v) bad("<" + n.name + "> lacks the attribute '" + key + "'");
  return *v;
}

struct Params {  // the flattened <rgb> / <float> / <integer> / <string> / <boolean> children of an element
  std::map<std::string, std::array<float, 3>> rgb;
  std::map<std::string, float> f;
  std::map<std::string, int> i;
  std::map<std::string, std::string> s;
  const std::array<float, 3>* first_rgb = nullptr;
  const std::string* first_string_value = nullptr;
  std::vector<std::array<float, 3>> rgb_store;
  std::vector<std::string> str_store;
};
void read_params(const XmlNode& n, Params* p) {
  p->rgb_store.reserve(n.children.size());
  p->str_store.reserve(n.children.size());
  for (const auto& c : n.children) {
    if (c->name == "rgb") {
      const std::vector<float> v = parse_floats(need_attr(*c, "value"));
      if (v.size() < 3) bad("<rgb> needs three components");
      p->rgb_store.push_back({v[0], v[1], v[2]});
      if (!p->first_rgb) p->first_rgb = &p->rgb_store.back();
      p->rgb[snake_case(c->attr_or("name", ""))] = p->rgb_store.back();
    } else if (c->name == "float") {
      p->f[snake_case(c->attr_or("name", ""))] = parse_float(need_attr(*c, "value"));
    } else if (c->name == "integer") {
      p->i[snake_case(c->attr_or("name", ""))] = std::atoi(need_attr(*c, "value").c_str());
    } else if (c->name == "string" || c->name == "boolean") {
      p->str_store.push_back(need_attr(*c, "value"));
      if (c->name == "string" && !p->first_string_value) p->first_string_value = &p->str_store.back();
      p->s[snake_case(c->attr_or("name", ""))] = p->str_store.back();
    }
  }
}
template <class M>
const typename M::mapped_type& need(const M& m, const char* key, const char* where) {
  auto it = m.find(key);
  if (it == m.end()) bad(std::string(where) + " lacks the parameter '" + key + "'");
  return it->second;
}

M4 parse_transform(const XmlNode& owner) {  // mod transform, common/importer/mitsuba.rs:274-295
  const XmlNode* t = owner.child("transform");
  if (!t) bad("<" + owner.name + "> lacks a <transform>");
  const XmlNode* m = t->child("matrix");
  if (!m) bad("<transform> lacks a <matrix>");
  const std::vector<float> v = parse_floats(need_attr(*m, "value"));
  if (v.size() != 16) bad("<matrix> needs 16 values");
  M4 out;
  for (int k = 0; k < 16; ++k) out.m[k] = v[k];  // from_row_slice
  return out;
}

std::string sibling_path(const std::string& scene_path, const std::string& file) {
  const size_t slash = scene_path.find_last_of('/');
  return slash == std::string::npos ? file : scene_path.substr(0, slash + 1) + file;
}

struct Importer {
  const std::string& path;
  const ImportOptions& opt;
  SceneBuilder& b;
  std::map<std::string, int> materials;

  // texture_from_mitsuba, pathtracer/importer/mitsuba.rs:24-68
  int texture(const XmlNode& t) {
    const std::string& type = need_attr(t, "type");
    Params p;
    read_params(t, &p);
    if (type == "checkerboard") {
      const auto& c0 = need(p.rgb, "color0", "checkerboard");
      const auto& c1 = need(p.rgb, "color1", "checkerboard");
      return b.add_checker_texture(3, c0.data(), c1.data(), need(p.f, "uscale", "checkerboard"), need(p.f, "vscale", "checkerboard"),
                                   need(p.f, "uoffset", "checkerboard"), need(p.f, "voffset", "checkerboard"));
    }
    if (type == "bitmap") {
      const ImageU8 img = load_image(sibling_path(path, need(p.s, "filename", "bitmap")));
      if (img.channels != 3) bad("unsupported image format for texture");  // only DynamicImage::ImageRgb8
      std::vector<float> f(img.data.size());
      for (size_t k = 0; k < f.size(); ++k) f[k] = 1.0f * inverse_gamma_correct((float)img.data[k] / 255.0f);
      return b.add_image_texture(3, f.data(), img.width, img.height, PTRS_WRAP_REPEAT, 1.f, -1.f, 0.f, 0.f);
    }
    bad("unknown texture type '" + type + "'");
  }
  // texture_with_defaults, :70-83
  int texture_or(const XmlNode* tex, const std::array<float, 3>* rgb) {
    if (tex) return texture(*tex);
    if (rgb) return b.add_constant_texture(3, (*rgb)[0], (*rgb)[1], (*rgb)[2]);
    return b.add_constant_texture(3, 1.f, 1.f, 1.f);
  }
  static const std::array<float, 3>* find(const Params& p, const char* key) {
    auto it = p.rgb.find(key);
    return it == p.rgb.end() ? nullptr : &it->second;
  }
  int spectrum_const(const std::array<float, 3>& v) { return b.add_constant_texture(3, v[0], v[1], v[2]); }

  // material_from_bsdf, :85-181
  int material(const XmlNode& n) {
    const std::string& type = need_attr(n, "type");
    if (type == "twosided") {
      const XmlNode* inner = n.child("bsdf");
      if (!inner) bad("twosided bsdf without a nested bsdf");
      return material(*inner);
    }
    Params p;
    read_params(n, &p);
    const XmlNode* tex = n.child("texture");
    if (type == "diffuse") {
      static const std::array<float, 3> one = {1.f, 1.f, 1.f};  // default_rgb_one
      return b.add_matte(texture_or(tex, p.first_rgb ? p.first_rgb : &one));
    }
    if (type == "conductor") {
      if (p.first_string_value) {
        if (*p.first_string_value == "none") return b.add_mirror();
        bad("other material values not supported yet!");
      }
      const int rough = b.add_constant_texture(1, 0.001f);
      return b.add_metal(spectrum_const(need(p.rgb, "eta", "conductor")), spectrum_const(need(p.rgb, "k", "conductor")),
                         texture_or(tex, find(p, "specular_reflectance")), rough, rough, false);
    }
    if (type == "roughconductor") {
      const int rough = b.add_constant_texture(1, need(p.f, "alpha", "roughconductor"));
      return b.add_metal(spectrum_const(need(p.rgb, "eta", "roughconductor")), spectrum_const(need(p.rgb, "k", "roughconductor")),
                         texture_or(tex, find(p, "specular_reflectance")), rough, rough, false);
    }
    if (type == "dielectric")
      return b.add_glass(b.add_constant_texture(3, 1.f, 1.f, 1.f), b.add_constant_texture(3, 1.f, 1.f, 1.f),
                         b.add_constant_texture(1, need(p.f, "int_ior", "dielectric")));
    if (type == "plastic" || type == "roughplastic") {
      const float eta = need(p.f, "int_ior", type.c_str());
      const float r0 = ((eta - 1.0f) * (eta - 1.0f)) / ((eta + 1.0f) * (eta + 1.0f));  // schlick_r0_from_eta, material/mod.rs:93-95
      const float alpha = type == "plastic" ? 0.001f : need(p.f, "alpha", "roughplastic");
      const int kd = texture_or(tex, find(p, "diffuse_reflectance"));
      const int ks = b.add_constant_texture(3, r0, r0, r0);
      const int nu = b.add_constant_texture(1, alpha), nv = b.add_constant_texture(1, alpha);
      return b.add_substrate(kd, ks, nu, nv, false);
    }
    bad("unknown bsdf type '" + type + "'");
  }

  // load_obj, common/importer/mitsuba.rs:81-151 (wavefront_obj: one object, one geometry, position / normal /
  // texture indices must coincide)
  MeshInput load_obj(const std::string& file) {
    const std::vector<uint8_t> bytes = read_file(sibling_path(path, file));
    std::istringstream in(std::string(bytes.begin(), bytes.end()));
    MeshInput m;
    std::string line;
    int objects = 0;
    while (std::getline(in, line)) {
      std::istringstream ls(line);
      std::string tag;
      if (!(ls >> tag)) continue;
      if (tag == "v") {
        double x, y, z;
        ls >> x >> y >> z;
        m.pos.insert(m.pos.end(), {(float)x, (float)y, (float)z});
      } else if (tag == "vn") {
        double x, y, z;
        ls >> x >> y >> z;
        m.normal.insert(m.normal.end(), {(float)x, (float)y, (float)z});
      } else if (tag == "vt") {
        double u = 0, v = 0;
        ls >> u >> v;
        m.uv.insert(m.uv.end(), {(float)u, (float)v});
      } else if (tag == "o") {
        if (++objects > 1) bad("only supporting one object right now!");
      } else if (tag == "f") {
        std::vector<uint32_t> face;
        std::string vert;
        while (ls >> vert) {
          long idx[3] = {0, 0, 0};
          int k = 0;
          size_t start = 0;
          for (size_t c = 0; c <= vert.size() && k < 3; ++c)
            if (c == vert.size() || vert[c] == '/') {
              if (c > start) idx[k] = std::atol(vert.substr(start, c - start).c_str());
              ++k;
              start = c + 1;
            }
          if (idx[0] <= 0) bad("OBJ: relative / missing vertex indices are not supported");
          if (idx[2] == 0 || idx[2] != idx[0]) bad("OBJ: normal index must equal the position index");
          if (idx[1] != 0 && idx[1] != idx[0]) bad("OBJ: texture index must equal the position index");
          face.push_back((uint32_t)(idx[0] - 1));
        }
        if (face.size() < 3) bad("OBJ: face with fewer than three vertices");
        for (size_t k = 1; k + 1 < face.size(); ++k) m.indices.insert(m.indices.end(), {face[0], face[k], face[k + 1]});
      }
    }
    return m;
  }

  // parse_shape, pathtracer/importer/mitsuba.rs:183-330
  void shape(const XmlNode& n) {
    const std::string& type = need_attr(n, "type");
    MeshInput m;
    if (type == "rectangle") {
      m = gen_rectangle();
      m.obj_to_world = parse_transform(n);
    } else if (type == "cube") {
      m = gen_cube();
      m.obj_to_world = parse_transform(n);
    } else if (type == "sphere") {
      const XmlNode* pt = n.child("point");
      const XmlNode* rad = n.child("float");
      if (!pt || !rad) bad("sphere needs <point> and <float radius>");
      const float c[3] = {parse_float(need_attr(*pt, "x")), parse_float(need_attr(*pt, "y")), parse_float(need_attr(*pt, "z"))};
      const float r = parse_float(need_attr(*rad, "value"));
      m = gen_sphere_uv(10, 10, false);
      for (size_t k = 0; k < m.pos.size(); ++k) m.pos[k] = m.pos[k] * r + c[k % 3];  // Similarity3(center, 0, radius) * p
    } else if (type == "obj") {
      Params p;
      read_params(n, &p);
      m = load_obj(need(p.s, "filename", "obj shape"));
      auto fn = p.s.find("face_normals");
      if (fn != p.s.end() && fn->second == "true") m.normal.clear();
      m.obj_to_world = parse_transform(n);
    } else {
      bad("unknown shape type '" + type + "'");
    }
    if (const XmlNode* ref = n.child("ref")) {
      auto it = materials.find(need_attr(*ref, "id"));
      if (it == materials.end()) bad("shape references the unknown bsdf '" + need_attr(*ref, "id") + "'");
      m.material = it->second;
    } else if (const XmlNode* embedded = n.child("bsdf")) {
      m.material = material(*embedded);
    } else {
      bad("either ref exists or embedded bsdf exists");
    }
    if (const XmlNode* em = n.child("emitter")) {
      if (need_attr(*em, "type") == "area") {
        Params p;
        read_params(*em, &p);
        if (!p.first_rgb) bad("area emitter without <rgb>");
        m.ke_tex = spectrum_const(*p.first_rgb);
      }
    }
    b.add_mesh(m);
  }

  void env_light(const M4& light_to_world, const std::string& hdr_path) {
    if (hdr_path.empty()) {
      const std::vector<float> sky = synth_sky(1024, 512, 1);
      b.add_infinite_light(light_to_world, sky.data(), 1024, 512);
      return;
    }
    const ImageF32 img = load_hdr(hdr_path);
    b.add_infinite_light(light_to_world, img.data.data(), img.width, img.height);
  }

  PtrsCamera run() {
    const std::vector<uint8_t> bytes = read_file(path);
    const std::string src(bytes.begin(), bytes.end());
    XmlParser parser(src);
    const std::unique_ptr<XmlNode> root = parser.parse_document();
    if (root->name != "scene") bad("root element is not <scene>");
    const XmlNode* sensor = root->child("sensor");
    if (!sensor) bad("no <sensor>");
    Params sp;
    read_params(*sensor, &sp);
    const XmlNode* film = sensor->child("film");
    if (!film) bad("<sensor> lacks a <film>");
    Params fp;
    read_params(*film, &fp);
    const PtrsCamera cam = mitsuba_camera(parse_transform(*sensor), need(sp.f, "fov", "sensor"), need(fp.i, "width", "film"),
                                          need(fp.i, "height", "film"), opt.res_w, opt.res_h);
    // the reference keeps these in a HashMap (arbitrary order); material numbering has no effect on the image
    for (const XmlNode* bs : root->all("bsdf")) materials[need_attr(*bs, "id")] = material(*bs);
    for (const XmlNode* sh : root->all("shape")) shape(*sh);
    const M4 env_to_world = mitsuba_env_light_to_world();  // importer/mitsuba.rs:365-372
    for (const XmlNode* em : root->all("emitter")) {
      const std::string& type = need_attr(*em, "type");
      if (type == "envmap") {
        Params p;
        read_params(*em, &p);
        if (!p.first_string_value) bad("envmap emitter without a filename");
        env_light(parse_transform(*em) * env_to_world, sibling_path(path, *p.first_string_value));
      } else if (type == "sunsky") {
        env_light(env_to_world, opt.sunsky_hdr);
      }  // standalone area emitters are an error log, point emitters are ignored (:375-378)
    }
    return cam;
  }
};

}  // namespace

PtrsCamera import_mitsuba(const std::string& path, const ImportOptions& opt, SceneBuilder& b) {
  Importer imp{path, opt, b, {}};
  return imp.run();
}

PtrsCamera import_scene(const std::string& path, const ImportOptions& opt, SceneBuilder& b) {
  const size_t dot = path.find_last_of('.');
  const std::string ext = dot == std::string::npos ? "" : path.substr(dot + 1);
  if (ext == "gltf" || ext == "glb") return import_gltf(path, opt, b);
  if (ext == "xml") return import_mitsuba(path, opt, b);
  throw std::runtime_error("unsupported format!");
}

}  // namespace ptrs_host

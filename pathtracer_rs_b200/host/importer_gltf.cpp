// glTF 2.0 (.gltf + external / data-URI buffers, .glb) -> SceneBuilder.  Follows src/common/importer/gltf.rs
// (camera) and src/pathtracer/importer/gltf.rs (materials :171-289, meshes :291-383, node walk and lights :385-503,
// scene assembly :505-584).  The reference reads the file through the `gltf` crate (1.0, not in the checkout:
// accessor decoding, TRS / matrix node transforms, KHR_lights_punctual, KHR_materials_transmission / _ior) and
// decodes images with `image`; here PNG and JPEG images are decoded (image_io.cpp, jpeg_decode.cpp).
#include <algorithm>
#include <array>
#include <cmath>
#include <cstring>
#include <map>
#include <stdexcept>

#include "image_io.hpp"
#include "importers.hpp"
#include "json_lite.hpp"
#include "procedural.hpp"

namespace ptrs_host {
namespace {

[[noreturn]] void bad(const std::string& what) { throw std::runtime_error("glTF import: " + what); }

std::vector<uint8_t> base64_decode(const std::string& s, size_t from) {
  std::vector<uint8_t> out;
  uint32_t acc = 0;
  int bits = 0;
  for (size_t i = from; i < s.size(); ++i) {
    const char c = s[i];
    int v;
    if (c >= 'A' && c <= 'Z') v = c - 'A';
    else if (c >= 'a' && c <= 'z') v = c - 'a' + 26;
    else if (c >= '0' && c <= '9') v = c - '0' + 52;
    else if (c == '+' || c == '-') v = 62;
    else if (c == '/' || c == '_') v = 63;
    else continue;  // '=', whitespace
    acc = (acc << 6) | (uint32_t)v;
    bits += 6;
    if (bits >= 8) {
      bits -= 8;
      out.push_back((uint8_t)(acc >> bits));
    }
  }
  return out;
}

std::string dir_of(const std::string& path) {
  const size_t slash = path.find_last_of('/');
  return slash == std::string::npos ? std::string() : path.substr(0, slash + 1);
}

// UnitQuaternion::to_rotation_matrix (nalgebra 0.32) into the upper 3x3 of a homogeneous matrix
M4 quat_to_m4(float i, float j, float k, float w) {
  const float ww = w * w, ii = i * i, jj = j * j, kk = k * k;
  const float ij = i * j * 2.0f, wk = w * k * 2.0f, wj = w * j * 2.0f, ik = i * k * 2.0f, jk = j * k * 2.0f, wi = w * i * 2.0f;
  M4 m = M4::identity();
  m.at(0, 0) = ww + ii - jj - kk;
  m.at(0, 1) = ij - wk;
  m.at(0, 2) = wj + ik;
  m.at(1, 0) = wk + ij;
  m.at(1, 1) = ww - ii + jj - kk;
  m.at(1, 2) = jk - wi;
  m.at(2, 0) = ik - wj;
  m.at(2, 1) = wi + jk;
  m.at(2, 2) = ww - ii - jj + kk;
  return m;
}

struct Trs {
  float t[3] = {0, 0, 0}, r[4] = {0, 0, 0, 1}, s[3] = {1, 1, 1};
};

// gltf::scene::Transform::decomposed for the matrix form (columns m[c][r]): translation = last column, scale =
// column lengths (z signed by the determinant), rotation = quaternion of the normalised basis
Trs decompose(const float m[16]) {  // column-major as stored in the file
  Trs o;
  o.t[0] = m[12];
  o.t[1] = m[13];
  o.t[2] = m[14];
  float x[3] = {m[0], m[1], m[2]}, y[3] = {m[4], m[5], m[6]}, z[3] = {m[8], m[9], m[10]};
  auto len = [](const float v[3]) { return std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); };
  const float det = x[0] * (y[1] * z[2] - z[1] * y[2]) - y[0] * (x[1] * z[2] - z[1] * x[2]) + z[0] * (x[1] * y[2] - y[1] * x[2]);
  const float sx = len(x), sy = len(y), sz = (det < 0.f ? -1.f : 1.f) * len(z);
  o.s[0] = sx;
  o.s[1] = sy;
  o.s[2] = sz;
  for (int k = 0; k < 3; ++k) {
    x[k] *= 1.0f / sx;
    y[k] *= 1.0f / sy;
    z[k] *= 1.0f / sz;
  }
  // Quaternion::from(Matrix3) (cgmath-style): x, y, z are the columns
  const float trace = x[0] + y[1] + z[2];
  float qs, qx, qy, qz;
  if (trace >= 0.f) {
    float s = std::sqrt(1.0f + trace);
    qs = 0.5f * s;
    s = 0.5f / s;
    qx = (y[2] - z[1]) * s;
    qy = (z[0] - x[2]) * s;
    qz = (x[1] - y[0]) * s;
  } else if (x[0] > y[1] && x[0] > z[2]) {
    float s = std::sqrt((x[0] - y[1] - z[2]) + 1.0f);
    qx = 0.5f * s;
    s = 0.5f / s;
    qy = (y[0] + x[1]) * s;
    qz = (x[2] + z[0]) * s;
    qs = (y[2] - z[1]) * s;
  } else if (y[1] > z[2]) {
    float s = std::sqrt((y[1] - x[0] - z[2]) + 1.0f);
    qy = 0.5f * s;
    s = 0.5f / s;
    qz = (z[1] + y[2]) * s;
    qx = (y[0] + x[1]) * s;
    qs = (z[0] - x[2]) * s;
  } else {
    float s = std::sqrt((z[2] - x[0] - y[1]) + 1.0f);
    qz = 0.5f * s;
    s = 0.5f / s;
    qx = (x[2] + z[0]) * s;
    qy = (z[1] + y[2]) * s;
    qs = (x[1] - y[0]) * s;
  }
  o.r[0] = qx;
  o.r[1] = qy;
  o.r[2] = qz;
  o.r[3] = qs;
  return o;
}

struct Importer {
  const std::string& path;
  const ImportOptions& opt;
  SceneBuilder& b;
  Json doc;
  std::vector<std::vector<uint8_t>> buffers;
  std::vector<uint8_t> glb_bin;
  bool has_glb_bin = false;
  std::map<long, ImageU8> image_cache;
  std::vector<int> materials;  // [0] = default matte, [i + 1] = document material i
  struct Deferred {
    int kind;  // 0 directional
    M4 xf;
    float color[3];
  };
  std::vector<Deferred> deferred;
  float wb_min[3] = {INFINITY, INFINITY, INFINITY}, wb_max[3] = {-INFINITY, -INFINITY, -INFINITY};
  bool any_geometry = false;

  const Json& list(const char* key, long i) const {
    const Json* a = doc.get(key);
    if (!a || i < 0 || (size_t)i >= a->size()) bad(std::string("index out of range in '") + key + "'");
    return a->at((size_t)i);
  }

  // ---- buffers / accessors ----------------------------------------------------------------------------
  void load_buffers() {
    const Json* bl = doc.get("buffers");
    for (size_t i = 0; bl && i < bl->size(); ++i) {
      const Json& bj = bl->at(i);
      const Json* uri = bj.get("uri");
      if (!uri) {
        if (!has_glb_bin || i != 0) bad("buffer without a uri outside a .glb");
        buffers.push_back(glb_bin);
      } else if (uri->str.rfind("data:", 0) == 0) {
        const size_t comma = uri->str.find(',');
        if (comma == std::string::npos) bad("malformed data uri");
        buffers.push_back(base64_decode(uri->str, comma + 1));
      } else {
        buffers.push_back(read_file(dir_of(path) + uri->str));
      }
      if (buffers.back().size() < (size_t)bj.number_or("byteLength", 0)) bad("buffer shorter than its byteLength");
    }
  }
  struct View {
    const uint8_t* p;
    size_t stride, count;
    int comp_type, n_comp;
    bool normalized;
  };
  View accessor(long idx) const {
    const Json& a = list("accessors", idx);
    if (a.get("sparse")) bad("sparse accessors are not supported");
    const long bv_i = a.index_or("bufferView", -1);
    if (bv_i < 0) bad("accessor without a bufferView");
    const Json& bv = list("bufferViews", bv_i);
    const long buf = bv.index_or("buffer", -1);
    if (buf < 0 || (size_t)buf >= buffers.size()) bad("bufferView names an unknown buffer");
    View v;
    v.comp_type = (int)a.index_or("componentType", 0);
    const std::string type = a.string_or("type", "");
    v.n_comp = type == "SCALAR" ? 1 : type == "VEC2" ? 2 : type == "VEC3" ? 3 : type == "VEC4" ? 4 : 0;
    if (!v.n_comp) bad("unsupported accessor type '" + type + "'");
    size_t csz;
    switch (v.comp_type) {
      case 5120: case 5121: csz = 1; break;
      case 5122: case 5123: csz = 2; break;
      case 5125: case 5126: csz = 4; break;
      default: bad("unknown componentType");
    }
    const size_t elem = csz * (size_t)v.n_comp;
    v.stride = (size_t)bv.index_or("byteStride", 0);
    if (v.stride == 0) v.stride = elem;
    v.count = (size_t)a.index_or("count", 0);
    const size_t off = (size_t)bv.index_or("byteOffset", 0) + (size_t)a.index_or("byteOffset", 0);
    if (v.count && off + (v.count - 1) * v.stride + elem > buffers[(size_t)buf].size()) bad("accessor runs past the end of its buffer");
    v.p = buffers[(size_t)buf].data() + off;
    v.normalized = a.get("normalized") && a.at("normalized").b;
    return v;
  }
  static float component(const View& v, size_t i, int c) {
    const uint8_t* e = v.p + i * v.stride;
    switch (v.comp_type) {
      case 5126: {
        float f;
        std::memcpy(&f, e + 4 * c, 4);
        return f;
      }
      case 5121: return v.normalized ? (float)e[c] / 255.0f : (float)e[c];
      case 5123: {
        uint16_t u;
        std::memcpy(&u, e + 2 * c, 2);
        return v.normalized ? (float)u / 65535.0f : (float)u;
      }
      case 5120: return v.normalized ? std::fmax((float)(int8_t)e[c] / 127.0f, -1.0f) : (float)(int8_t)e[c];
      case 5122: {
        int16_t s;
        std::memcpy(&s, e + 2 * c, 2);
        return v.normalized ? std::fmax((float)s / 32767.0f, -1.0f) : (float)s;
      }
      default: {
        uint32_t u;
        std::memcpy(&u, e + 4 * c, 4);
        return (float)u;
      }
    }
  }
  std::vector<float> read_floats(long idx, int take) const {  // first `take` components of every element
    const View v = accessor(idx);
    if (v.n_comp < take) bad("accessor has too few components");
    std::vector<float> out(v.count * (size_t)take);
    for (size_t i = 0; i < v.count; ++i)
      for (int c = 0; c < take; ++c) out[i * take + c] = component(v, i, c);
    return out;
  }
  std::vector<uint32_t> read_indices(long idx) const {
    const View v = accessor(idx);
    std::vector<uint32_t> out(v.count);
    for (size_t i = 0; i < v.count; ++i) {
      const uint8_t* e = v.p + i * v.stride;
      if (v.comp_type == 5121) out[i] = e[0];
      else if (v.comp_type == 5123) {
        uint16_t u;
        std::memcpy(&u, e, 2);
        out[i] = u;
      } else if (v.comp_type == 5125) {
        std::memcpy(&out[i], e, 4);
      } else {
        bad("index accessor must be u8 / u16 / u32");
      }
    }
    return out;
  }

  // ---- images / textures -------------------------------------------------------------------------------
  const ImageU8& image(long idx) {
    auto it = image_cache.find(idx);
    if (it != image_cache.end()) return it->second;
    const Json& ij = list("images", idx);
    std::vector<uint8_t> bytes;
    if (const Json* uri = ij.get("uri")) {
      if (uri->str.rfind("data:", 0) == 0) {
        const size_t comma = uri->str.find(',');
        if (comma == std::string::npos) bad("malformed data uri");
        bytes = base64_decode(uri->str, comma + 1);
      } else {
        bytes = read_file(dir_of(path) + uri->str);
      }
    } else {
      const Json& bv = list("bufferViews", ij.index_or("bufferView", -1));
      const long buf = bv.index_or("buffer", -1);
      if (buf < 0 || (size_t)buf >= buffers.size()) bad("image bufferView names an unknown buffer");
      const size_t off = (size_t)bv.index_or("byteOffset", 0), len = (size_t)bv.index_or("byteLength", 0);
      if (off + len > buffers[(size_t)buf].size()) bad("image bufferView runs past the end of its buffer");
      bytes.assign(buffers[(size_t)buf].begin() + off, buffers[(size_t)buf].begin() + off + len);
    }
    return image_cache.emplace(idx, decode_image(bytes.data(), bytes.size())).first->second;
  }
  struct TexRef {
    long image;
    int wrap;
    float scale;  // normalTexture.scale
  };
  TexRef tex_ref(const Json& info) {
    const Json& t = list("textures", info.index_or("index", -1));
    TexRef r;
    r.image = t.index_or("source", -1);
    r.scale = (float)info.number_or("scale", 1.0);
    long ws = 10497, wt = 10497;
    const long si = t.index_or("sampler", -1);
    if (si >= 0) {
      const Json& s = list("samplers", si);
      ws = s.index_or("wrapS", 10497);
      wt = s.index_or("wrapT", 10497);
    }
    if (ws != wt) bad("sampler wrapS != wrapT");  // assert_eq!(sampler.wrap_s(), sampler.wrap_t())
    r.wrap = ws == 33071 ? PTRS_WRAP_CLAMP : PTRS_WRAP_REPEAT;  // wrap_mode_from_gtlf: mirrored repeat -> repeat
    return r;
  }
  // color_texture_from_gltf (:39-96): factor * inverse_gamma(pixel / 255); -1 when the format is not RGB / RGBA
  int color_texture(const Json& info, const float factor[3]) {
    const TexRef r = tex_ref(info);
    const ImageU8& img = image(r.image);
    if (img.channels != 3 && img.channels != 4) return -1;
    std::vector<float> f((size_t)img.width * img.height * 3);
    for (size_t p = 0; p < (size_t)img.width * img.height; ++p)
      for (int c = 0; c < 3; ++c) f[3 * p + c] = factor[c] * inverse_gamma_correct((float)img.data[p * img.channels + c] / 255.0f);
    return b.add_image_texture(3, f.data(), img.width, img.height, r.wrap, 1.f, 1.f, 0.f, 0.f);
  }
  // ImageTexture::<f32>::new over one channel: scale * (pixel / 255)
  int channel_texture(const TexRef& r, int channel, float scale) {
    const ImageU8& img = image(r.image);
    std::vector<float> f((size_t)img.width * img.height);
    for (size_t p = 0; p < f.size(); ++p) f[p] = scale * ((float)img.data[p * img.channels + channel] / 255.0f);
    return b.add_image_texture(1, f.data(), img.width, img.height, r.wrap, 1.f, 1.f, 0.f, 0.f);
  }

  // ---- materials (material_from_gltf, :171-289) -----------------------------------------------------------
  int material(const Json& m) {
    static const Json empty;
    const Json* pbr_p = m.get("pbrMetallicRoughness");
    const Json& pbr = pbr_p ? *pbr_p : empty;
    float base[4] = {1, 1, 1, 1};
    if (const Json* f = pbr.get("baseColorFactor"))
      for (int c = 0; c < 4 && c < (int)f->size(); ++c) base[c] = (float)f->at((size_t)c).num;
    const float color_factor[3] = {inverse_gamma_correct(base[0]), inverse_gamma_correct(base[1]), inverse_gamma_correct(base[2])};  // from_slice_4(.., true)
    int color_tex = b.add_constant_texture(3, color_factor[0], color_factor[1], color_factor[2]);
    if (const Json* info = pbr.get("baseColorTexture")) {
      const int t = color_texture(*info, color_factor);
      if (t >= 0) color_tex = t;
    }
    int normal_tex = -1;
    if (const Json* info = m.get("normalTexture")) {
      const TexRef r = tex_ref(*info);
      const ImageU8& img = image(r.image);
      if (img.channels != 3) bad("normal maps must be 8-bit RGB images");
      std::vector<float> f(img.data.size());
      for (size_t p = 0; p < (size_t)img.width * img.height; ++p) {  // NormalMap::new, texture.rs:152-178
        f[3 * p] = r.scale * ((float)img.data[3 * p] / 127.5f - 1.0f);
        f[3 * p + 1] = r.scale * ((float)img.data[3 * p + 1] / 127.5f - 1.0f);
        f[3 * p + 2] = (float)img.data[3 * p + 2] / 127.5f - 1.0f;
      }
      normal_tex = b.add_image_texture(3, f.data(), img.width, img.height, r.wrap, 1.f, 1.f, 0.f, 0.f);
    }
    float transmission = 0.f, ior = 1.5f;
    if (const Json* ext = m.get("extensions")) {
      if (const Json* t = ext->get("KHR_materials_transmission")) transmission = (float)t->number_or("transmissionFactor", 0.0);
      if (const Json* i = ext->get("KHR_materials_ior")) ior = (float)i->number_or("ior", 1.5);
    }
    auto with_normal = [&](int mat) {
      if (normal_tex >= 0) b.set_normal_map(mat, normal_tex);
      return mat;
    };
    const int one = b.add_constant_texture(3, 1.f, 1.f, 1.f);
    if (transmission == 1.0f) return with_normal(b.add_glass(one, b.add_constant_texture(3, 1.f, 1.f, 1.f), b.add_constant_texture(1, ior)));
    const float alpha = base[3];
    if (m.string_or("alphaMode", "OPAQUE") == "BLEND" && alpha < 1.0f) {
      const int kt = b.add_constant_texture(3, 1.0f - alpha * color_factor[0], 1.0f - alpha * color_factor[1], 1.0f - alpha * color_factor[2]);
      return with_normal(b.add_glass(one, kt, b.add_constant_texture(1, 1.33f)));
    }
    const float metallic = (float)pbr.number_or("metallicFactor", 1.0), roughness = (float)pbr.number_or("roughnessFactor", 1.0);
    if (metallic == 1.0f && roughness == 0.0f) return b.add_mirror();
    int metallic_tex = b.add_constant_texture(1, metallic), roughness_tex = b.add_constant_texture(1, roughness);
    if (const Json* info = pbr.get("metallicRoughnessTexture")) {
      const TexRef r = tex_ref(*info);
      const ImageU8& img = image(r.image);
      if (img.channels == 3 || img.channels == 4) {  // metallic = blue, roughness = green (:118-150)
        metallic_tex = channel_texture(r, 2, metallic);
        roughness_tex = channel_texture(r, 1, roughness);
      }
    }
    return with_normal(b.add_disney(color_tex, metallic_tex, b.add_constant_texture(1, ior), roughness_tex));
  }

  // ---- node walk (populate_scene, :385-503) -------------------------------------------------------------------
  static M4 node_transform(const Json& n) {  // trans_from_gltf: t * r * s of the decomposed transform
    Trs trs;
    if (const Json* mj = n.get("matrix")) {
      float m[16];
      for (int k = 0; k < 16; ++k) m[k] = (float)mj->at((size_t)k).num;
      trs = decompose(m);
    } else {
      if (const Json* t = n.get("translation"))
        for (int k = 0; k < 3; ++k) trs.t[k] = (float)t->at((size_t)k).num;
      if (const Json* r = n.get("rotation"))
        for (int k = 0; k < 4; ++k) trs.r[k] = (float)r->at((size_t)k).num;
      if (const Json* s = n.get("scale"))
        for (int k = 0; k < 3; ++k) trs.s[k] = (float)s->at((size_t)k).num;
    }
    M4 T = M4::identity(), S = M4::identity();
    for (int k = 0; k < 3; ++k) {
      T.at(k, 3) = trs.t[k];
      S.at(k, k) = trs.s[k];
    }
    return T * quat_to_m4(trs.r[0], trs.r[1], trs.r[2], trs.r[3]) * S;
  }

  void primitive(const Json& prim, const M4& xf) {
    if (prim.index_or("mode", 4) != 4) bad("only triangle-list primitives are supported");
    const long mat_i = prim.index_or("material", -1);
    static const Json empty;
    const Json& mat = mat_i >= 0 ? list("materials", mat_i) : empty;
    MeshInput mesh;
    // emission: channel 0 of the factor for all three channels, times 10 (:398-405)
    float e0 = 0.f;
    if (const Json* ef = mat.get("emissiveFactor")) e0 = (float)ef->at(0).num;
    const float ke[3] = {10.0f * e0, 10.0f * e0, 10.0f * e0};
    const bool emissive = ke[0] != 0.0f;
    bool textured_ke = false;
    int ke_image_tex = -1;
    if (emissive) {
      mesh.ke_tex = b.add_constant_texture(3, ke[0], ke[1], ke[2]);
      if (const Json* info = mat.get("emissiveTexture")) {
        const int t = color_texture(*info, ke);
        if (t >= 0) {
          mesh.ke_tex = ke_image_tex = t;
          textured_ke = true;
        }
      }
    }
    // alpha mask from the base colour texture's alpha channel (:300-329)
    const Json* pbr = mat.get("pbrMetallicRoughness");
    if (pbr && pbr->get("baseColorTexture") && mat.string_or("alphaMode", "OPAQUE") == "MASK") {
      const TexRef r = tex_ref(pbr->at("baseColorTexture"));
      if (image(r.image).channels != 4) bad("alphaMode MASK needs an RGBA base colour texture");
      mesh.alpha_tex = channel_texture(r, 3, 1.0f);
    }
    const Json& attrs = prim.at("attributes");
    if (prim.index_or("indices", -1) < 0) bad("primitive without indices");  // read_indices().unwrap()
    mesh.indices = read_indices(prim.index_or("indices", -1));
    mesh.indices.resize(mesh.indices.size() / 3 * 3);  // chunks_exact(3)
    if (attrs.index_or("POSITION", -1) < 0) bad("primitive without POSITION");
    mesh.pos = read_floats(attrs.index_or("POSITION", -1), 3);
    if (attrs.index_or("NORMAL", -1) >= 0) mesh.normal = read_floats(attrs.index_or("NORMAL", -1), 3);
    if (attrs.index_or("TANGENT", -1) >= 0) mesh.tangent = read_floats(attrs.index_or("TANGENT", -1), 3);
    if (attrs.index_or("TEXCOORD_0", -1) >= 0) mesh.uv = read_floats(attrs.index_or("TEXCOORD_0", -1), 2);
    mesh.obj_to_world = xf;
    mesh.material = materials[(size_t)(mat_i + 1)];
    const size_t nt = mesh.indices.size() / 3, nv = mesh.pos.size() / 3;
    for (uint32_t idx : mesh.indices)
      if (idx >= nv) bad("vertex index out of range");
    for (size_t i = 0; i < mesh.indices.size(); ++i) {
      const uint32_t v = mesh.indices[i];
      const V3 p = xform_point(xf, v3(mesh.pos[3 * v], mesh.pos[3 * v + 1], mesh.pos[3 * v + 2]));
      for (int k = 0; k < 3; ++k) {
        wb_min[k] = std::fmin(wb_min[k], p[k]);
        wb_max[k] = std::fmax(wb_max[k], p[k]);
      }
      any_geometry = true;
    }
    if (textured_ke) {
      // a triangle becomes a light only if ke is non-black at one of 10 x 10 sample points (:424-448)
      const HostMipMap* mm = b.image_texture_mip(ke_image_tex);
      mesh.tri_emits.assign(nt, 0);
      for (size_t t = 0; t < nt && mm; ++t) {
        float uv[3][2] = {{0.f, 0.f}, {1.f, 0.f}, {1.f, 1.f}};  // Triangle::get_uvs default, shape.rs:34-48
        if (!mesh.uv.empty())
          for (int k = 0; k < 3; ++k) {
            uv[k][0] = mesh.uv[2 * mesh.indices[3 * t + k]];
            uv[k][1] = mesh.uv[2 * mesh.indices[3 * t + k] + 1];
          }
        for (int x = 0; x < 10 && !mesh.tri_emits[t]; ++x)
          for (int y = 0; y < 10; ++y) {
            const float u0 = (float)x * 0.1f, u1 = (float)y * 0.1f;
            const float su0 = std::sqrt(u0), b0 = 1.0f - su0, b1 = u1 * su0;  // uniform_sample_triangle, shape.rs:14-17
            const float s = b0 * uv[0][0] + b1 * uv[1][0] + (1.0f - b0 - b1) * uv[2][0];
            const float tt = b0 * uv[0][1] + b1 * uv[1][1] + (1.0f - b0 - b1) * uv[2][1];
            float c[3];
            mm->lookup_width(s, tt, 0.0f, c);
            if (c[0] != 0.f || c[1] != 0.f || c[2] != 0.f) {
              mesh.tri_emits[t] = 1;
              break;
            }
          }
      }
    }
    b.add_mesh(mesh);
  }

  void node(const M4& parent, long idx) {
    const Json& n = list("nodes", idx);
    const M4 xf = parent * node_transform(n);
    if (n.index_or("mesh", -1) >= 0) {
      const Json& mesh = list("meshes", n.index_or("mesh", -1));
      const Json& prims = mesh.at("primitives");
      for (size_t p = 0; p < prims.size(); ++p) primitive(prims.at(p), xf);
    }
    if (const Json* ext = n.get("extensions"))
      if (const Json* lp = ext->get("KHR_lights_punctual")) {
        const Json* dext = doc.get("extensions");
        const Json* dl = dext ? dext->get("KHR_lights_punctual") : nullptr;
        const Json* lights = dl ? dl->get("lights") : nullptr;
        const long li = lp->index_or("light", -1);
        if (!lights || li < 0 || (size_t)li >= lights->size()) bad("node references an unknown punctual light");
        const Json& l = lights->at((size_t)li);
        float c0 = 1.f;
        if (const Json* col = l.get("color")) c0 = (float)col->at(0).num;
        const float v = (float)l.number_or("intensity", 1.0) * c0;  // channel 0 for all three (:466-470)
        const float color[3] = {v, v, v};
        if (l.string_or("type", "") == "directional") {
          Deferred d{0, xf, {v, v, v}};
          deferred.push_back(d);
        } else {  // point; spot lights are treated as point lights (:485-491)
          b.add_point_light(xf, color);
        }
      }
    if (const Json* ch = n.get("children"))
      for (size_t c = 0; c < ch->size(); ++c) node(xf, (long)ch->at(c).num);
  }

  // ---- camera (common/importer/gltf.rs) ------------------------------------------------------------------
  bool find_camera(const M4& parent, long idx, PtrsCamera* out) {
    const Json& n = list("nodes", idx);
    const M4 xf = parent * node_transform(n);
    const long ci = n.index_or("camera", -1);
    if (ci >= 0) {
      const Json& cam = list("cameras", ci);
      if (cam.string_or("type", "") == "perspective") {
        const Json& p = cam.at("perspective");
        float rot[9];
        for (int r = 0; r < 3; ++r)
          for (int c = 0; c < 3; ++c) rot[3 * r + c] = xf.at(r, c);
        // na::try_convert::<Transform3, Isometry3>: the linear part must be a rotation
        for (int c = 0; c < 3; ++c) {
          const float l2 = rot[c] * rot[c] + rot[3 + c] * rot[3 + c] + rot[6 + c] * rot[6 + c];
          if (std::fabs(l2 - 1.0f) > 1e-4f) bad("camera node transform is not an isometry");
        }
        float q[4];
        quat_from_matrix(rot, q);
        const float t[3] = {xf.at(0, 3), xf.at(1, 3), xf.at(2, 3)};
        *out = make_camera(q, t, (float)opt.res_w / (float)opt.res_h, (float)p.number_or("yfov", 1.0), (float)p.number_or("znear", 0.01),
                           (float)p.number_or("zfar", 1000.0), opt.res_w, opt.res_h);
        return true;
      }
    }
    // `for child in children { return find_camera(child) }`: only the first child is ever searched
    if (const Json* ch = n.get("children"))
      if (ch->size() > 0) return find_camera(xf, (long)ch->at(0).num, out);
    return false;
  }
  PtrsCamera default_camera() const {  // get_default_camera: look from world_bound.p_max at the origin
    const V3 eye = any_geometry ? v3(wb_max[0], wb_max[1], wb_max[2]) : v3(1, 1, 1);
    const V3 z = normalize(eye - v3(0, 0, 0)), x = normalize(cross(v3(0, 1, 0), z)), y = cross(z, x);
    const float rot[9] = {x.x, y.x, z.x, x.y, y.y, z.y, x.z, y.z, z.z};
    float q[4];
    quat_from_matrix(rot, q);
    const float t[3] = {eye.x, eye.y, eye.z};
    const float rx = (float)opt.res_w, ry = (float)opt.res_h;
    return make_camera(q, t, rx / ry, 1.57079632679489661923f * (ry / rx), 0.01f, 1000.0f, opt.res_w, opt.res_h);
  }

  PtrsCamera run() {
    std::vector<uint8_t> bytes = read_file(path);
    const char* json = reinterpret_cast<const char*>(bytes.data());
    size_t json_len = bytes.size();
    if (bytes.size() >= 12 && std::memcmp(bytes.data(), "glTF", 4) == 0) {  // binary container
      size_t off = 12;
      bool have_json = false;
      while (off + 8 <= bytes.size()) {
        uint32_t len, type;
        std::memcpy(&len, bytes.data() + off, 4);
        std::memcpy(&type, bytes.data() + off + 4, 4);
        if (off + 8 + (size_t)len > bytes.size()) bad("truncated .glb chunk");
        if (type == 0x4E4F534Au && !have_json) {
          json = reinterpret_cast<const char*>(bytes.data() + off + 8);
          json_len = len;
          have_json = true;
        } else if (type == 0x004E4942u && !has_glb_bin) {
          glb_bin.assign(bytes.begin() + off + 8, bytes.begin() + off + 8 + len);
          has_glb_bin = true;
        }
        off += 8 + (size_t)len;
      }
      if (!have_json) bad(".glb without a JSON chunk");
    }
    doc = JsonParser(json, json_len).parse_document();
    load_buffers();
    materials.push_back(b.add_matte(b.add_constant_texture(3, 1.f, 1.f, 1.f)));  // default_material
    if (const Json* ml = doc.get("materials"))
      for (size_t i = 0; i < ml->size(); ++i) materials.push_back(material(ml->at(i)));
    const Json* scenes = doc.get("scenes");
    for (size_t s = 0; scenes && s < scenes->size(); ++s)
      if (const Json* nodes = scenes->at(s).get("nodes"))
        for (size_t k = 0; k < nodes->size(); ++k) node(M4::identity(), (long)nodes->at(k).num);
    // lights that need the world bound come after the others (:548-575)
    for (const Deferred& d : deferred) {
      const float w[3] = {0.f, 0.f, -1.f};
      b.add_directional_light(d.xf, d.color, w);
    }
    if (opt.default_lights) {
      // UnitQuaternion::from_euler_angles(-pi/2, 0, 0): the env map is z-up, the scene y-up
      const float hr = -1.57079632679489661923f * 0.5f;
      const float sr = std::sin(hr), cr = std::cos(hr), sp = 0.f, cp = 1.f, sy = 0.f, cy = 1.f;
      const M4 l2w = quat_to_m4(sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy, cr * cp * cy + sr * sp * sy);
      if (opt.sunsky_hdr.empty()) {
        const std::vector<float> sky = synth_sky(1024, 512, 1);
        b.add_infinite_light(l2w, sky.data(), 1024, 512);
      } else {
        const ImageF32 img = load_hdr(opt.sunsky_hdr);
        b.add_infinite_light(l2w, img.data.data(), img.width, img.height);
      }
    }
    PtrsCamera cam = default_camera();
    bool found = false;
    for (size_t s = 0; scenes && s < scenes->size() && !found; ++s)
      if (const Json* nodes = scenes->at(s).get("nodes"))
        for (size_t k = 0; k < nodes->size() && !found; ++k) found = find_camera(M4::identity(), (long)nodes->at(k).num, &cam);
    return cam;
  }
};

}  // namespace

PtrsCamera import_gltf(const std::string& path, const ImportOptions& opt, SceneBuilder& b) {
  Importer imp{path, opt, b};
  return imp.run();
}

}  // namespace ptrs_host

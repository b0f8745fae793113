//! src/pathtracer/gpu/tables.rs — interning tables that turn a `RenderScene`'s object graph (Arc'd meshes, materials,
//! boxed textures, trait-object lights) into the flat pools of `PtrsSceneDesc` (include/ptrs_b200.h).
//!
//! Identity is by address: the importers share one `Arc<TriangleMesh>` / `Arc<Material>` among all triangles of a mesh
//! (importer/mitsuba.rs:334-362, importer/gltf.rs:430-520), so a pointer map recovers the mesh / material tables.
//! Uses the exporters of `reference_additions.rs`.  Written against the reference's types; NOT compiled here (the
//! pathtracer-b200 image has no Rust toolchain) — tests/test_rust_shim.py checks it textually against b200.rs and ffi.rs.
use super::b200::FlatScene;
use super::ffi;
use crate::common::WrapMode;
use crate::pathtracer::light::{LightExport, SyncLight};
use crate::pathtracer::material::Material;
use crate::pathtracer::shape::{Triangle, TriangleMesh};
use crate::pathtracer::texture::{SyncTexture, TextureExport};
use std::collections::HashMap;
use std::sync::Arc;

fn addr<T: ?Sized>(p: &T) -> usize {
    p as *const T as *const u8 as usize
}

#[derive(Default)]
pub struct Tables {
    mesh_ids: HashMap<usize, (i32, u32)>, // Arc<TriangleMesh> address -> (mesh id, first vertex in the pools)
    material_ids: HashMap<usize, i32>,    // Arc<Material> address -> material id
    light_ids: HashMap<usize, i32>,       // DiffuseAreaLight address -> index in RenderScene::lights
    area_prims: HashMap<usize, i32>,      // Triangle address -> BVH-ordered primitive index (filled by note_prim)
    pos: Vec<f32>,
    normal: Vec<f32>,
    tangent: Vec<f32>,
    uv: Vec<f32>,
    any_normal: bool,
    any_tangent: bool,
    any_uv: bool,
    meshes: Vec<ffi::PtrsMesh>,
    materials: Vec<ffi::PtrsMaterial>,
    textures: Vec<ffi::PtrsTexture>,
    mipmaps: Vec<ffi::PtrsMipMap>,
    texels: Vec<f32>,
    lights: Vec<Arc<dyn SyncLight>>,
    infinite: Vec<Arc<dyn SyncLight>>,
}

impl Tables {
    /// Vertex pools in `RenderScene::meshes` order; positions, normals and tangents are already in world space
    /// (shape.rs:602-612).  A pool a mesh lacks is zero-filled so that one global vertex index addresses all four.
    pub fn new(meshes: &[Arc<TriangleMesh>], lights: &[Arc<dyn SyncLight>], infinite: &[Arc<dyn SyncLight>]) -> Self {
        let mut t = Tables::default();
        t.lights = lights.to_vec();
        t.infinite = infinite.to_vec();
        for (i, l) in lights.iter().enumerate() {
            if let LightExport::Area { .. } = l.export() {
                t.light_ids.insert(addr(l.as_ref()), i as i32);
            }
        }
        for mesh in meshes {
            let base = (t.pos.len() / 3) as u32;
            let n = mesh.pos.len();
            for p in &mesh.pos {
                t.pos.extend_from_slice(&[p.x, p.y, p.z]);
            }
            let mut flags = 0u32;
            let mut pool = |dst: &mut Vec<f32>, src: &[na::Vector3<f32>], any: &mut bool, bit: u32| {
                if src.len() == n {
                    for v in src {
                        dst.extend_from_slice(&[v.x, v.y, v.z]);
                    }
                    *any = true;
                    flags |= bit;
                } else {
                    dst.resize(dst.len() + 3 * n, 0.0);
                }
            };
            pool(&mut t.normal, &mesh.normal, &mut t.any_normal, ffi::PTRS_MESH_HAS_NORMAL as u32);
            pool(&mut t.tangent, &mesh.s, &mut t.any_tangent, ffi::PTRS_MESH_HAS_TANGENT as u32);
            if mesh.uv.len() == n {
                for v in &mesh.uv {
                    t.uv.extend_from_slice(&[v.x, v.y]);
                }
                t.any_uv = true;
                flags |= ffi::PTRS_MESH_HAS_UV as u32;
            } else {
                t.uv.resize(t.uv.len() + 2 * n, 0.0);
            }
            let alpha_tex = match &mesh.alpha_mask {
                Some(a) => t.float_texture(a.as_ref()),
                None => -1,
            };
            let id = t.meshes.len() as i32;
            t.meshes.push(ffi::PtrsMesh { flags, alpha_tex });
            t.mesh_ids.insert(addr(mesh.as_ref()), (id, base));
        }
        t
    }

    pub fn vertex_base(&self, mesh: &Arc<TriangleMesh>) -> u32 {
        self.mesh_ids[&addr(mesh.as_ref())].1
    }
    pub fn mesh_id(&self, mesh: &Arc<TriangleMesh>) -> i32 {
        self.mesh_ids[&addr(mesh.as_ref())].0
    }
    /// Index in `RenderScene::lights` of the DiffuseAreaLight a primitive carries (primitive.rs:21-25).
    pub fn light_id(&self, light: &crate::pathtracer::light::DiffuseAreaLight) -> i32 {
        self.light_ids[&addr(light)]
    }
    /// Remembers where a triangle ended up in BVH order: `PtrsLight::prim` of its area light.
    pub fn note_prim(&mut self, tri: &Arc<Triangle>, prim_index: usize) {
        self.area_prims.insert(addr(tri.as_ref()), prim_index as i32);
    }

    pub fn material_id(&mut self, material: &Arc<Material>) -> i32 {
        let key = addr(material.as_ref());
        if let Some(id) = self.material_ids.get(&key) {
            return *id;
        }
        let e = material.export();
        let mut m = ffi::PtrsMaterial { type_: e.kind, normal_map: -1, tex: [-1; 5], remap_roughness: e.remap_roughness as i32 };
        for (slot, t) in e.spectrum_tex {
            m.tex[slot] = self.spectrum_texture(t);
        }
        for (slot, t) in e.float_tex {
            m.tex[slot] = self.float_texture(t);
        }
        if let Some(n) = e.normal_map {
            m.normal_map = self.push_texture(n.export(), 3);
        }
        let id = self.materials.len() as i32;
        self.materials.push(m);
        self.material_ids.insert(key, id);
        id
    }

    fn spectrum_texture(&mut self, t: &dyn SyncTexture<crate::common::spectrum::Spectrum>) -> i32 {
        self.push_texture(t.export(), 3)
    }
    fn float_texture(&mut self, t: &dyn SyncTexture<f32>) -> i32 {
        self.push_texture(t.export(), 1)
    }

    fn push_texture(&mut self, e: TextureExport, channels: i32) -> i32 {
        let mut out = ffi::PtrsTexture { type_: 0, channels, v1: [0.0; 3], v2: [0.0; 3], su: 1.0, sv: 1.0, du: 0.0, dv: 0.0, mip: -1, pad: 0 };
        match e {
            TextureExport::Constant(v) => {
                out.type_ = ffi::PTRS_TEX_CONSTANT;
                out.v1 = v;
            }
            TextureExport::Checker { v1, v2, map } => {
                out.type_ = ffi::PTRS_TEX_CHECKER;
                out.v1 = v1;
                out.v2 = v2;
                let (su, sv, du, dv) = map.params();
                out.su = su; out.sv = sv; out.du = du; out.dv = dv;
            }
            TextureExport::Image { levels, channels: c, wrap, map } => {
                out.type_ = ffi::PTRS_TEX_IMAGE;
                let (su, sv, du, dv) = map.params();
                out.su = su; out.sv = sv; out.du = du; out.dv = dv;
                out.mip = self.push_mip(&levels, c as i32, wrap);
            }
        }
        self.textures.push(out);
        self.textures.len() as i32 - 1
    }

    /// The pyramid as MIPMap::new built it (texture.rs:279-405).  Handing over `&levels[..1]` instead lets the
    /// library build the remaining levels on the device (include/ptrs_b200.h, PtrsMipMap).
    fn push_mip(&mut self, levels: &[(usize, usize, Vec<f32>)], channels: i32, wrap: WrapMode) -> i32 {
        let mut m: ffi::PtrsMipMap = unsafe { std::mem::zeroed() };
        m.channels = channels;
        m.wrap = match wrap {
            WrapMode::Repeat => ffi::PTRS_WRAP_REPEAT,
            WrapMode::Black => ffi::PTRS_WRAP_BLACK,
            WrapMode::Clamp => ffi::PTRS_WRAP_CLAMP,
        };
        m.n_levels = levels.len() as i32;
        for (l, (w, h, texels)) in levels.iter().enumerate() {
            m.width[l] = *w as i32;
            m.height[l] = *h as i32;
            m.level_offset[l] = self.texels.len() as u64;
            self.texels.extend_from_slice(texels);
        }
        self.mipmaps.push(m);
        self.mipmaps.len() as i32 - 1
    }

    /// Moves the pools into `flat` and appends the light tables (RenderScene::lights / infinite_lights order).
    pub fn write_pools(mut self, flat: &mut FlatScene) {
        let lights = std::mem::take(&mut self.lights);
        let infinite = std::mem::take(&mut self.infinite);
        for l in &lights {
            let mut out: ffi::PtrsLight = unsafe { std::mem::zeroed() };
            out.prim = -1;
            out.ke_tex = -1;
            out.env = -1;
            match l.export() {
                LightExport::Point { p, i } => {
                    out.type_ = ffi::PTRS_LIGHT_POINT;
                    out.pos = [p.x, p.y, p.z];
                    out.color = [i.r(), i.g(), i.b()];
                }
                LightExport::Directional { w, l, world_center, world_radius } => {
                    out.type_ = ffi::PTRS_LIGHT_DIRECTIONAL;
                    out.pos = [w.x, w.y, w.z];
                    out.color = [l.r(), l.g(), l.b()];
                    out.world_center = [world_center.x, world_center.y, world_center.z];
                    out.world_radius = world_radius;
                }
                LightExport::Area { shape, ke, area } => {
                    out.type_ = ffi::PTRS_LIGHT_AREA;
                    out.prim = self.area_prims[&addr(shape.as_ref())];
                    out.ke_tex = self.spectrum_texture(ke);
                    out.area = area;
                }
                LightExport::Infinite { light_to_world, world_to_light, world_center, world_radius, l_map, distribution } => {
                    out.type_ = ffi::PTRS_LIGHT_INFINITE;
                    out.world_center = [world_center.x, world_center.y, world_center.z];
                    out.world_radius = world_radius;
                    out.env = flat.push_env(&mut self, light_to_world, world_to_light, l_map, distribution);
                }
            }
            flat.lights.push(out);
        }
        // infinite_lights holds ids into `lights` (mod.rs:84-89: the same Arcs appear in both lists, importer/mitsuba.rs:397-398)
        for inf in &infinite {
            if let Some(i) = lights.iter().position(|l| addr(l.as_ref()) == addr(inf.as_ref())) {
                flat.infinite_lights.push(i as i32);
            }
        }
        flat.pos = self.pos;
        flat.normal = if self.any_normal { self.normal } else { Vec::new() };
        flat.tangent = if self.any_tangent { self.tangent } else { Vec::new() };
        flat.uv = if self.any_uv { self.uv } else { Vec::new() };
        flat.meshes = self.meshes;
        flat.materials = self.materials;
        flat.textures = self.textures;
        flat.mipmaps = self.mipmaps;
        flat.texels = self.texels;
    }

    /// MIP pyramid of an environment map (Spectrum texels, WrapMode::Repeat: light.rs:349-371).
    pub fn push_env_mip(&mut self, l_map: &crate::pathtracer::texture::MIPMap<crate::common::spectrum::Spectrum>) -> i32 {
        let levels: Vec<(usize, usize, Vec<f32>)> = l_map
            .levels()
            .iter()
            .map(|m| {
                let mut texels = Vec::with_capacity(m.nrows() * m.ncols() * 3);
                for row in 0..m.nrows() {
                    for col in 0..m.ncols() {
                        let s = m[(row, col)];
                        texels.extend_from_slice(&[s.r(), s.g(), s.b()]);
                    }
                }
                (m.ncols(), m.nrows(), texels)
            })
            .collect();
        self.push_mip(&levels, 3, WrapMode::Repeat)
    }
}

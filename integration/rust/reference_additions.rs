//! The additions pathtracer-rs itself needs so that `gpu/b200.rs` + `gpu/tables.rs` can read a `RenderScene`.
//!
//! The reference keeps the fields the GPU path has to copy private (accelerator.rs:97-100, primitive.rs:21-25,
//! shape.rs:20-25, texture.rs, material/*.rs, light.rs) and hands textures, materials and lights around as trait
//! objects, so a flat description cannot be produced from outside those modules.  Each block below is appended to the
//! file named above it (inside that module private fields are visible); nothing existing is modified.  Together with
//! `mod b200; mod ffi; mod tables;` in src/pathtracer/gpu/mod.rs and the feature-gated call in
//! `PathIntegrator::render` this is the whole patch.  Not compiled here: the image has no rustc.

// ---------------------------------------------------------------------------------------------------------------------
// src/pathtracer/accelerator.rs
// ---------------------------------------------------------------------------------------------------------------------
impl BVH {
    /// The flattened tree (accelerator.rs:348-357) and the primitives in `ordered_prims` order: exactly the two
    /// arrays `intersect` walks (accelerator.rs:359-417).
    pub fn export_flat(&self) -> (&[LinearBVHNode], &[Arc<dyn SyncPrimitive>]) {
        (&self.nodes, &self.primitives)
    }
}
impl LinearBVHNode {
    /// `primitives_offset` / `second_child_offset`: the union at accelerator.rs:85-88 read as its `u32`.
    pub fn offset(&self) -> u32 {
        unsafe { self.offset.primitives_offset }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// src/pathtracer/primitive.rs  (the trait gains one method; GeometricPrimitive is its only implementor)
// ---------------------------------------------------------------------------------------------------------------------
pub trait Primitive {
    // ... existing methods ...
    /// The triangle and the shared material, for exporters (every primitive of the reference is a
    /// `GeometricPrimitive` over a `Triangle`).
    fn get_shape(&self) -> &Arc<Triangle>;
    fn get_material_arc(&self) -> &Arc<Material>;
}
impl Primitive for GeometricPrimitive {
    // ... existing methods ...
    fn get_shape(&self) -> &Arc<Triangle> {
        &self.shape
    }
    fn get_material_arc(&self) -> &Arc<Material> {
        &self.material
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// src/pathtracer/shape.rs
// ---------------------------------------------------------------------------------------------------------------------
impl Triangle {
    pub fn mesh(&self) -> &Arc<TriangleMesh> {
        &self.mesh
    }
    pub fn indices(&self) -> &[u32; 3] {
        &self.indices
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// src/pathtracer/texture.rs — what a texture is made of, as plain data
// ---------------------------------------------------------------------------------------------------------------------
pub enum TextureExport<'a> {
    Constant([f32; 3]),
    Checker { v1: [f32; 3], v2: [f32; 3], map: &'a UVMap },
    /// MIP pyramid levels, finest first, row-major, `channels` floats per texel (texture.rs:238-243)
    Image { levels: Vec<(usize, usize, Vec<f32>)>, channels: usize, wrap: WrapMode, map: &'a UVMap },
}
pub trait Texture<T> {
    fn evaluate(&self, it: &SurfaceMediumInteraction) -> T;
    /// Plain-data view for exporters.
    fn export(&self) -> TextureExport;
}
pub trait Channels: Copy {
    fn rgb(self) -> [f32; 3];
    const N: usize;
}
impl Channels for f32 {
    fn rgb(self) -> [f32; 3] {
        [self, self, self]
    }
    const N: usize = 1;
}
impl Channels for Spectrum {
    fn rgb(self) -> [f32; 3] {
        [self.r(), self.g(), self.b()]
    }
    const N: usize = 3;
}
impl Channels for na::Vector3<f32> {
    fn rgb(self) -> [f32; 3] {
        [self.x, self.y, self.z]
    }
    const N: usize = 3;
}
// in `impl<T: Copy + Channels> Texture<T> for ConstantTexture<T>`:
//     fn export(&self) -> TextureExport { TextureExport::Constant(self.value.rgb()) }
// in `impl<T: Copy + Channels> Texture<T> for CheckerTexture<T>`:
//     fn export(&self) -> TextureExport { TextureExport::Checker { v1: self.v1.rgb(), v2: self.v2.rgb(), map: &self.mapping } }
// in `impl<T: na::Scalar + num::Zero + Channels> Texture<T> for ImageTexture<T>`:
//     fn export(&self) -> TextureExport {
//         let levels = self.mip_map.pyramid.iter().map(|m| {
//             let mut texels = Vec::with_capacity(m.nrows() * m.ncols() * T::N);
//             for row in 0..m.nrows() { for col in 0..m.ncols() { texels.extend_from_slice(&m[(row, col)].rgb()[..T::N]); } }
//             (m.ncols(), m.nrows(), texels)
//         }).collect();
//         TextureExport::Image { levels, channels: T::N, wrap: self.mip_map.wrap_mode, map: &self.mapping }
//     }
impl<T: na::Scalar + num::Zero> MIPMap<T> {
    pub fn levels(&self) -> &[na::DMatrix<T>] {
        &self.pyramid
    }
}
impl UVMap {
    pub fn params(&self) -> (f32, f32, f32, f32) {
        (self.su, self.sv, self.du, self.dv)
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// src/pathtracer/material/mod.rs — (type, parameter textures in the PtrsMaterial::tex order, remap flag, normal map)
// ---------------------------------------------------------------------------------------------------------------------
pub struct MaterialExport<'a> {
    pub kind: i32, // PTRS_MAT_*
    pub spectrum_tex: Vec<(usize, &'a dyn SyncTexture<Spectrum>)>, // (slot in PtrsMaterial::tex, texture)
    pub float_tex: Vec<(usize, &'a dyn SyncTexture<f32>)>,
    pub remap_roughness: bool,
    pub normal_map: Option<&'a dyn SyncTexture<na::Vector3<f32>>>,
}
impl Material {
    pub fn export(&self) -> MaterialExport {
        let plain = |kind, s: Vec<(usize, &dyn SyncTexture<Spectrum>)>, f: Vec<(usize, &dyn SyncTexture<f32>)>, remap| MaterialExport {
            kind,
            spectrum_tex: s,
            float_tex: f,
            remap_roughness: remap,
            normal_map: None,
        };
        match self {
            Material::Matte(m) => plain(0, vec![(0, m.kd.as_ref())], vec![], false),
            Material::Mirror(_) => plain(1, vec![], vec![], false),
            Material::Glass(m) => plain(2, vec![(0, m.kr.as_ref()), (1, m.kt.as_ref())], vec![(2, m.index.as_ref())], false),
            // MetalMaterial resolves `roughness` into u / v when they are absent (metal.rs:62-73)
            Material::Metal(m) => plain(
                3,
                vec![(0, m.eta.as_ref()), (1, m.k.as_ref()), (2, m.r.as_ref())],
                vec![(3, m.u_roughness.as_deref().or(m.roughness.as_deref()).unwrap()), (4, m.v_roughness.as_deref().or(m.roughness.as_deref()).unwrap())],
                m.remap_roughness,
            ),
            Material::Substrate(m) => plain(4, vec![(0, m.kd.as_ref()), (1, m.ks.as_ref())], vec![(2, m.nu.as_ref()), (3, m.nv.as_ref())], m.remap_roughness),
            Material::Disney(m) => plain(5, vec![(0, m.color.as_ref())], vec![(1, m.metallic.as_ref()), (2, m.eta.as_ref()), (3, m.roughness.as_ref())], false),
            Material::Normal(m) => {
                let mut inner = m.material.export();
                inner.normal_map = Some(m.normal_map.as_ref());
                inner
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// src/pathtracer/light.rs — the Light trait gains `export`
// ---------------------------------------------------------------------------------------------------------------------
pub enum LightExport<'a> {
    Point { p: na::Point3<f32>, i: Spectrum },
    Directional { w: na::Vector3<f32>, l: Spectrum, world_center: na::Point3<f32>, world_radius: f32 },
    Area { shape: &'a Arc<Triangle>, ke: &'a dyn SyncTexture<Spectrum>, area: f32 },
    Infinite {
        light_to_world: &'a na::Projective3<f32>,
        world_to_light: &'a na::Projective3<f32>,
        world_center: na::Point3<f32>,
        world_radius: f32,
        l_map: &'a MIPMap<Spectrum>,
        distribution: &'a Distribution2D,
    },
}
// `fn export(&self) -> LightExport;` in `trait Light`, and in the four impls:
//   PointLight:        LightExport::Point { p: self.p_light, i: self.i }
//   DirectionalLight:  LightExport::Directional { w: self.w_light, l: self.l, world_center: self.world_center, world_radius: self.world_radius }
//   DiffuseAreaLight:  LightExport::Area { shape: &self.shape, ke: self.ke.as_ref(), area: self.area }
//   InfiniteAreaLight: LightExport::Infinite { light_to_world: &self.light_to_world, world_to_light: &self.world_to_light,
//                                              world_center: self.world_center, world_radius: self.world_radius,
//                                              l_map: &self.l_map, distribution: &self.distribution }

// ---------------------------------------------------------------------------------------------------------------------
// src/pathtracer/sampling.rs
// ---------------------------------------------------------------------------------------------------------------------
impl Distribution1D {
    pub fn parts(&self) -> (&[f32], &[f32], f32) {
        (&self.func, &self.cdf, self.func_int)
    }
}
impl Distribution2D {
    pub fn parts(&self) -> (&[Box<Distribution1D>], &Distribution1D) {
        (&self.p_conditional_v, &self.p_marginal)
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// src/pathtracer/mod.rs
// ---------------------------------------------------------------------------------------------------------------------
impl RenderScene {
    pub fn bvh(&self) -> &accelerator::BVH {
        &self.scene
    }
}

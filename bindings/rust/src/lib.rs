// backends/b200/src/lib.rs — `impl rfw_backend::Backend for B200Backend` over the C ABI of include/rfwb200.h.
// Source only: no Rust toolchain exists in the build image, so this file has never been compiled.
use rfw::prelude::*;
use std::os::raw::{c_char, c_int, c_void};

// the reference's sampler tables, moved over unchanged from backends/gpu-rt/src/blue_noise.rs (data + create_blue_noise_buffer)
mod blue_noise;

#[repr(C)]
struct RfwMeshData3D {
    triangles: *const RTTriangle, num_triangles: u32,
    vertices: *const Vertex3D, num_vertices: u32,
    ranges: *const VertexMesh, num_ranges: u32,
    skin_data: *const JointData, num_skin_data: u32,
    flags: u32, bounds: Aabb,
}
#[repr(C)]
struct RfwInstancesData3D { matrices: *const f32, skin_ids: *const i32, flags: *const u32, num_instances: u32, local_aabb: Aabb }
#[repr(C)]
struct RfwTextureData { width: u32, height: u32, mip_levels: u32, bytes: *const u8, num_bytes: u64, format: u32 }
#[repr(C)]
struct RfwSkinData { inverse_bind_matrices: *const f32, joint_matrices: *const f32, num_joints: u32 }
#[repr(C)]
#[derive(Default)]
struct RfwB200Config { device: i32, width: u32, height: u32, max_depth: u32, clamp_value: f32, tile_size: u32, rank: u32, world: u32, sky: [f32; 3], reserved: [u32; 8] }

extern "C" {
    fn rfwb200_create(cfg: *const RfwB200Config, out: *mut *mut c_void) -> c_int;
    fn rfwb200_destroy(h: *mut c_void);
    fn rfwb200_set_3d_mesh(h: *mut c_void, id: u32, data: *const RfwMeshData3D) -> c_int;
    fn rfwb200_unload_3d_meshes(h: *mut c_void, ids: *const u32, n: u32) -> c_int;
    fn rfwb200_set_3d_instances(h: *mut c_void, mesh: u32, data: *const RfwInstancesData3D) -> c_int;
    fn rfwb200_set_materials(h: *mut c_void, m: *const DeviceMaterial, n: u32, changed: *const u32) -> c_int;
    fn rfwb200_set_textures(h: *mut c_void, t: *const RfwTextureData, n: u32, changed: *const u32) -> c_int;
    fn rfwb200_set_skybox(h: *mut c_void, t: *const RfwTextureData) -> c_int;
    fn rfwb200_set_skins(h: *mut c_void, s: *const RfwSkinData, n: u32, changed: *const u32) -> c_int;
    fn rfwb200_set_point_lights(h: *mut c_void, l: *const PointLight, n: u32, changed: *const u32) -> c_int;
    fn rfwb200_set_spot_lights(h: *mut c_void, l: *const SpotLight, n: u32, changed: *const u32) -> c_int;
    fn rfwb200_set_area_lights(h: *mut c_void, l: *const AreaLight, n: u32, changed: *const u32) -> c_int;
    fn rfwb200_set_directional_lights(h: *mut c_void, l: *const DirectionalLight, n: u32, changed: *const u32) -> c_int;
    fn rfwb200_synchronize(h: *mut c_void) -> c_int;
    fn rfwb200_render(h: *mut c_void, view: *const CameraView3D, mode: u32) -> c_int;
    fn rfwb200_resize(h: *mut c_void, w: u32, hgt: u32, scale: f64) -> c_int;
    fn rfwb200_set_blue_noise(h: *mut c_void, table: *const u32, n: u32) -> c_int;
    // multi-GPU (one process per GPU; the accumulator gather over NCCL lives inside the library)
    pub fn rfwb200_comm_unique_id(out_id: *mut u8) -> c_int;
    pub fn rfwb200_comm_init(h: *mut c_void, unique_id: *const u8, rank: u32, world: u32) -> c_int;
    pub fn rfwb200_comm_destroy(h: *mut c_void) -> c_int;
    pub fn rfwb200_gather_image(h: *mut c_void, root: u32, d_image: *mut f32) -> c_int;
    pub fn rfwb200_render_gather(h: *mut c_void, view: *const CameraView3D, spp: u32, depth: u32, root: u32, d_image: *mut f32) -> c_int;
    fn rfwb200_last_error() -> *const c_char;
}

pub struct B200Backend { handle: *mut c_void }
unsafe impl Send for B200Backend {}
unsafe impl Sync for B200Backend {}

fn check(rc: c_int) {
    if rc != 0 {
        let msg = unsafe { std::ffi::CStr::from_ptr(rfwb200_last_error()) }.to_string_lossy().into_owned();
        panic!("rfwb200: {}", msg); // precedent: backends/metal/src/lib.rs:53
    }
}

impl FromWindowHandle for B200Backend {
    fn init<W: HasRawWindowHandle>(_w: &W, width: u32, height: u32, _scale: f64) -> Result<Box<Self>, Box<dyn std::error::Error>> {
        let cfg = RfwB200Config { width, height, max_depth: 3, clamp_value: 10.0, world: 1, ..Default::default() };
        let mut handle = std::ptr::null_mut();
        if unsafe { rfwb200_create(&cfg, &mut handle) } != 0 {
            let msg = unsafe { std::ffi::CStr::from_ptr(rfwb200_last_error()) }.to_string_lossy().into_owned();
            return Err(msg.into());
        }
        // the sampler tables of the first 256 samples: reference data (backends/gpu-rt/src/blue_noise.rs:40970-41004), handed over once
        let tables: Vec<u32> = blue_noise::create_blue_noise_buffer();
        check(unsafe { rfwb200_set_blue_noise(handle, tables.as_ptr(), tables.len() as u32) });
        Ok(Box::new(Self { handle }))
    }
}

impl Backend for B200Backend {
    fn set_2d_mesh(&mut self, _id: usize, _data: MeshData2D<'_>) {}
    fn set_2d_instances(&mut self, _mesh: usize, _instances: InstancesData2D<'_>) {}
    fn set_3d_mesh(&mut self, id: usize, d: MeshData3D<'_>) {
        let c = RfwMeshData3D {
            triangles: d.triangles.as_ptr(), num_triangles: d.triangles.len() as u32,
            vertices: d.vertices.as_ptr(), num_vertices: d.vertices.len() as u32,
            ranges: d.ranges.as_ptr(), num_ranges: d.ranges.len() as u32,
            skin_data: d.skin_data.as_ptr(), num_skin_data: d.skin_data.len() as u32,
            flags: d.flags.bits(), bounds: d.bounds,
        };
        check(unsafe { rfwb200_set_3d_mesh(self.handle, id as u32, &c) });
    }
    fn unload_3d_meshes(&mut self, ids: &[usize]) {
        let v: Vec<u32> = ids.iter().map(|i| *i as u32).collect();
        check(unsafe { rfwb200_unload_3d_meshes(self.handle, v.as_ptr(), v.len() as u32) });
    }
    fn set_3d_instances(&mut self, mesh: usize, i: InstancesData3D<'_>) {
        let c = RfwInstancesData3D {
            matrices: i.matrices.as_ptr() as *const f32, skin_ids: i.skin_ids.as_ptr() as *const i32,
            flags: i.flags.as_ptr() as *const u32, num_instances: i.len() as u32, local_aabb: i.local_aabb,
        };
        check(unsafe { rfwb200_set_3d_instances(self.handle, mesh as u32, &c) });
    }
    fn set_materials(&mut self, m: &[DeviceMaterial], _changed: &BitSlice) {
        check(unsafe { rfwb200_set_materials(self.handle, m.as_ptr(), m.len() as u32, std::ptr::null()) });
    }
    fn set_textures(&mut self, t: &[TextureData<'_>], changed: &BitSlice) {
        let c: Vec<RfwTextureData> = t.iter().map(|t| RfwTextureData {
            width: t.width, height: t.height, mip_levels: t.mip_levels, bytes: t.bytes.as_ptr(), num_bytes: t.bytes.len() as u64, format: t.format as u32,
        }).collect();
        let ch: Vec<u32> = (0..t.len()).map(|i| changed[i] as u32).collect();
        check(unsafe { rfwb200_set_textures(self.handle, c.as_ptr(), c.len() as u32, ch.as_ptr()) });
    }
    fn synchronize(&mut self) { check(unsafe { rfwb200_synchronize(self.handle) }); }
    fn render(&mut self, _v2: CameraView2D, v3: CameraView3D, mode: RenderMode) {
        check(unsafe { rfwb200_render(self.handle, &v3, mode as u32) });
    }
    fn resize(&mut self, size: (u32, u32), scale: f64) { check(unsafe { rfwb200_resize(self.handle, size.0, size.1, scale) }); }
    fn set_point_lights(&mut self, l: &[PointLight], _c: &BitSlice) { check(unsafe { rfwb200_set_point_lights(self.handle, l.as_ptr(), l.len() as u32, std::ptr::null()) }); }
    fn set_spot_lights(&mut self, l: &[SpotLight], _c: &BitSlice) { check(unsafe { rfwb200_set_spot_lights(self.handle, l.as_ptr(), l.len() as u32, std::ptr::null()) }); }
    fn set_area_lights(&mut self, l: &[AreaLight], _c: &BitSlice) { check(unsafe { rfwb200_set_area_lights(self.handle, l.as_ptr(), l.len() as u32, std::ptr::null()) }); }
    fn set_directional_lights(&mut self, l: &[DirectionalLight], _c: &BitSlice) { check(unsafe { rfwb200_set_directional_lights(self.handle, l.as_ptr(), l.len() as u32, std::ptr::null()) }); }
    fn set_skybox(&mut self, s: TextureData<'_>) {
        let c = RfwTextureData { width: s.width, height: s.height, mip_levels: s.mip_levels, bytes: s.bytes.as_ptr(), num_bytes: s.bytes.len() as u64, format: s.format as u32 };
        check(unsafe { rfwb200_set_skybox(self.handle, &c) });
    }
    fn set_skins(&mut self, s: &[SkinData<'_>], changed: &BitSlice) {
        let c: Vec<RfwSkinData> = s.iter().map(|s| RfwSkinData {
            inverse_bind_matrices: s.inverse_bind_matrices.as_ptr() as *const f32, joint_matrices: s.joint_matrices.as_ptr() as *const f32,
            num_joints: s.joint_matrices.len() as u32,
        }).collect();
        let ch: Vec<u32> = (0..s.len()).map(|i| changed[i] as u32).collect();
        check(unsafe { rfwb200_set_skins(self.handle, c.as_ptr(), c.len() as u32, ch.as_ptr()) });
    }
}

impl Drop for B200Backend { fn drop(&mut self) { unsafe { rfwb200_destroy(self.handle) } } }

#[cfg(test)]
mod tests {
    // the reference's ABI contract (backends/metal/src/lib.rs:270-348): Rust POD sizes == C header sizes
    use super::*;
    #[test]
    fn test_layout() {
        assert_eq!(std::mem::size_of::<RTTriangle>(), 176);
        assert_eq!(std::mem::size_of::<Vertex3D>(), 64);
        assert_eq!(std::mem::size_of::<DeviceMaterial>(), 96);
        assert_eq!(std::mem::size_of::<CameraView3D>(), 128);
        assert_eq!(std::mem::size_of::<AreaLight>(), 96);
        assert_eq!(std::mem::size_of::<SpotLight>(), 48);
        assert_eq!(std::mem::size_of::<PointLight>(), 32);
        assert_eq!(std::mem::size_of::<DirectionalLight>(), 32);
        assert_eq!(std::mem::size_of::<Aabb>(), 32);
        assert_eq!(std::mem::size_of::<VertexMesh>(), 48);
        assert_eq!(std::mem::size_of::<JointData>(), 32);
    }
}

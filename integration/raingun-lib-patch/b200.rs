//! raingun-lib/src/b200.rs (new) — `Scene::render_image` / `Scene::streaming_render` on the B200.
//!
//! Wiring (three edits to the reference, none of them in the CLI):
//!   raingun-lib/Cargo.toml   raingun-b200-sys = { path = "../raingun-b200-sys", optional = true }
//!                            [features] b200 = ["raingun-b200-sys"]
//!   raingun-lib/src/lib.rs   #[cfg(feature = "b200")] extern crate raingun_b200_sys; mod b200;
//!   raingun-lib/src/scene.rs in `render_image` (41-43) and `streaming_render` (45-51):
//!                            #[cfg(feature = "b200")] return b200::render_image(self, width, height);
//! SOURCE ONLY: not compiled here (no Rust toolchain); written against rustc 1.17-era syntax like the
//! reference (no `?` on Option, no dyn).
use std::ffi::CStr;
use std::os::raw::{c_int, c_void};
use std::sync::mpsc::Sender;

use image::{GenericImage, ImageBuffer, Rgba};
use raingun_b200_sys as sys;

use bodies::Body;
use color::Color;
use lights::Light;
use material::{Coloration, Surface};
use rendering::RenderedPixel;
use scene::Scene;

fn last_error() -> String {
    unsafe { CStr::from_ptr(sys::rg_last_error()).to_string_lossy().into_owned() }
}

/// An uploaded scene; freed on drop.
pub struct B200Scene(*mut sys::rg_scene);

impl Drop for B200Scene {
    fn drop(&mut self) {
        unsafe { sys::rg_scene_destroy(self.0) }
    }
}

/// Flattens `Vec<Body>` / `Vec<Light>` body by body, IN ORDER (the order is the tie-break of
/// `Scene::trace`, scene.rs:34-39), and uploads.  f32 fields are passed through as stored.
pub fn upload(scene: &Scene) -> B200Scene {
    let device = match ::std::env::var("RAINGUN_DEVICE") {
        Ok(v) => v.parse::<i32>().unwrap_or(sys::RG_DEVICE_ALL),
        Err(_) => sys::RG_DEVICE_ALL,
    };
    upload_on(scene, device)
}

/// The same on one named device (or `RG_DEVICE_ALL`).
pub fn upload_on(scene: &Scene, device: i32) -> B200Scene {
    let n = scene.bodies.len();
    let mut kind = vec![0u8; n];
    let mut geom = vec![0f64; 8 * n];
    let mut ckind = vec![0u8; n];
    let mut color = vec![0f32; 3 * n];
    let mut tex_id = vec![-1i32; n];
    let mut tex_off = vec![0f32; 2 * n];
    let mut albedo = vec![0f32; n];
    let mut skind = vec![0u8; n];
    let mut sparam = vec![0f32; 2 * n];
    let mut tex_pixels: Vec<Vec<u8>> = vec![];
    let mut textures: Vec<sys::rg_texture_desc> = vec![];

    for (i, body) in scene.bodies.iter().enumerate() {
        {
            let g = &mut geom[8 * i..8 * i + 8];
            match *body {
                Body::Sphere(ref s) => {                                   // bodies.rs:13-18
                    kind[i] = sys::RG_BODY_SPHERE;
                    g[..4].copy_from_slice(&[s.center.x, s.center.y, s.center.z, s.radius]);
                }
                Body::Plane(ref p) => {                                    // bodies.rs:20-25
                    kind[i] = sys::RG_BODY_PLANE;
                    g[..6].copy_from_slice(&[p.origin.x, p.origin.y, p.origin.z, p.normal.x, p.normal.y, p.normal.z]);
                }
                Body::Disk(ref d) => {                                     // bodies.rs:27-33
                    kind[i] = sys::RG_BODY_DISK;
                    g[..7].copy_from_slice(&[d.origin.x, d.origin.y, d.origin.z, d.normal.x, d.normal.y, d.normal.z, d.radius]);
                }
                Body::AABB(ref b) => {                                     // bodies.rs:35-39
                    kind[i] = sys::RG_BODY_AABB;
                    g[..6].copy_from_slice(&[b.bounds[0].x, b.bounds[0].y, b.bounds[0].z,
                                             b.bounds[1].x, b.bounds[1].y, b.bounds[1].z]);
                }
            }
        }
        let m = body.material();
        albedo[i] = m.albedo;
        match m.coloration {
            Coloration::Color(c) => {
                color[3 * i..3 * i + 3].copy_from_slice(&[c.red, c.green, c.blue]);
            }
            Coloration::Texture(ref t) => {
                ckind[i] = sys::RG_COLORATION_TEXTURE;
                tex_id[i] = textures.len() as i32;
                tex_off[2 * i] = t.x_offset;
                tex_off[2 * i + 1] = t.y_offset;
                let rgba = t.image.to_rgba();                      // what get_pixel() yields, material.rs:67
                textures.push(sys::rg_texture_desc { width: rgba.width(), height: rgba.height(), channels: 4,
                                                     reserved: 0, pixels: ::std::ptr::null() });
                tex_pixels.push(rgba.into_raw());
            }
        }
        match m.surface {
            Surface::Diffuse => {}
            Surface::Reflecting { reflectivity } => {
                skind[i] = sys::RG_SURFACE_REFLECTING;
                sparam[2 * i] = reflectivity;
            }
            Surface::Refractive { index, transparency } => {
                skind[i] = sys::RG_SURFACE_REFRACTIVE;
                sparam[2 * i] = index;
                sparam[2 * i + 1] = transparency;
            }
        }
    }
    for (t, px) in textures.iter_mut().zip(&tex_pixels) {
        t.pixels = px.as_ptr();
    }

    let nl = scene.lights.len();
    let mut lkind = vec![0u8; nl];
    let mut lvec = vec![0f64; 3 * nl];
    let mut lcolor = vec![0f32; 3 * nl];
    let mut lint = vec![0f32; nl];
    for (i, light) in scene.lights.iter().enumerate() {
        match *light {
            Light::Directional(ref d) => {                                 // lights.rs:8-13
                lkind[i] = sys::RG_LIGHT_DIRECTIONAL;
                lvec[3 * i..3 * i + 3].copy_from_slice(&[d.direction.x, d.direction.y, d.direction.z]);
                lcolor[3 * i..3 * i + 3].copy_from_slice(&[d.color.red, d.color.green, d.color.blue]);
                lint[i] = d.intensity;
            }
            Light::Spherical(ref s) => {                                   // lights.rs:15-20
                lkind[i] = sys::RG_LIGHT_SPHERICAL;
                lvec[3 * i..3 * i + 3].copy_from_slice(&[s.position.x, s.position.y, s.position.z]);
                lcolor[3 * i..3 * i + 3].copy_from_slice(&[s.color.red, s.color.green, s.color.blue]);
                lint[i] = s.intensity;
            }
        }
    }

    let desc = sys::rg_scene_desc {
        abi_version: sys::RG_ABI_VERSION,
        max_recursion_depth: scene.max_recursion_depth,
        fov: scene.fov,
        default_color: [scene.default_color.red, scene.default_color.green, scene.default_color.blue],
        n_bodies: n as u32,
        body_kind: kind.as_ptr(),
        body_geom: geom.as_ptr(),
        coloration_kind: ckind.as_ptr(),
        color: color.as_ptr(),
        texture_id: tex_id.as_ptr(),
        texture_offset: tex_off.as_ptr(),
        albedo: albedo.as_ptr(),
        surface_kind: skind.as_ptr(),
        surface_param: sparam.as_ptr(),
        n_lights: nl as u32,
        n_textures: textures.len() as u32,
        light_kind: lkind.as_ptr(),
        light_vec: lvec.as_ptr(),
        light_color: lcolor.as_ptr(),
        light_intensity: lint.as_ptr(),
        textures: if textures.is_empty() { ::std::ptr::null() } else { textures.as_ptr() },
    };
    let mut handle = ::std::ptr::null_mut();
    // RG_DEVICE_ALL: every visible GPU; the library splits a render into row tiles across them itself
    // (rayon's par_iter over the machine's cores, rendering.rs:27-35).
    let rc = unsafe { sys::rg_scene_create(&desc, device, &mut handle) };   // the library copies everything
    assert!(rc == sys::RG_OK, "raingun_b200: {}", last_error());         // the reference panics on failure too
    B200Scene(handle)
}

/// Scene::render_image (scene.rs:41-43 -> rendering.rs:24-38).
pub fn render_image(scene: &Scene, width: u32, height: u32) -> ImageBuffer<Rgba<u8>, Vec<u8>> {
    let gpu = upload(scene);
    let mut raw = vec![0u8; (width * height * 4) as usize];             // u32 arithmetic, as rendering.rs:27
    let rc = unsafe { sys::rg_render(gpu.0, width, height, raw.as_mut_ptr(), ::std::ptr::null_mut()) };
    assert!(rc == sys::RG_OK, "raingun_b200: {}", last_error());         // RG_E_PORTRAIT = ray.rs:42's assert
    ImageBuffer::from_raw(width, height, raw).unwrap()                   // rendering.rs:37
}

extern "C" fn on_rows(y0: u32, rows: u32, width: u32, rgb: *const f32, user: *mut c_void) -> c_int {
    let tx = unsafe { &*(user as *const Sender<RenderedPixel>) };
    let px = unsafe { ::std::slice::from_raw_parts(rgb, (rows * width * 3) as usize) };
    for (i, p) in px.chunks(3).enumerate() {
        let (x, y) = (i as u32 % width, y0 + i as u32 / width);
        // RenderedPixel.color is the UNQUANTISED f32 colour (rendering.rs:18-22,59-65): rg_render_stream_f32
        // delivers the reference's own f32 bits, so any consumer of the channel sees what it would have seen
        let color = Color { red: p[0], green: p[1], blue: p[2] };
        if tx.send(RenderedPixel { x: x, y: y, color: color }).is_err() {
            return 1;                                                    // closed channel: rendering.rs:53-54,67
        }
    }
    0
}

/// Scene::streaming_render (scene.rs:45-51 -> rendering.rs:40-69).
pub fn streaming_render(scene: &Scene, width: u32, height: u32, channel_tx: Sender<RenderedPixel>) {
    // the f32 stream is a single-device entry point: one GPU renders the bands
    let gpu = upload_on(scene, 0);
    let rc = unsafe {
        sys::rg_render_stream_f32(gpu.0, width, height, 0, on_rows, &channel_tx as *const _ as *mut c_void,
                                  ::std::ptr::null_mut())
    };
    assert!(rc == sys::RG_OK || rc == sys::RG_E_CANCELLED, "raingun_b200: {}", last_error());
}

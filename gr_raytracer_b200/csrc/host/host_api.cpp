// host_api.cpp -- C ABI of the host-side scene assembly and file I/O (include/grrt.h, second half).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>

#include "png.hpp"
#include "scene_build.hpp"

using namespace grrt_host;

struct GrrtHostScene {
    HostSceneData data;
};

namespace {
void set_err(char* err, size_t n, const std::string& msg) {
    if (!err || n == 0) return;
    std::snprintf(err, n, "%s", msg.c_str());
}
}  // namespace

extern "C" {

void grrt_cli_opts_default(GrrtCliOpts* o) {  // cli.rs:6-47
    if (!o) return;
    std::memset(o, 0, sizeof *o);
    o->width = 500;
    o->height = 500;
    o->step_size = 0.01;
    o->max_steps = 20000;
    o->max_radius = 15000.0;
    o->epsilon = 0.00001;
    o->camera_position[0] = 18.0;
    o->camera_position[1] = 0.0;
    o->camera_position[2] = 0.8;
    o->tone_mapping = GRRT_TONEMAP_REINHARD;
    o->sampling_mask_color[0] = 255;
    o->sampling_mask_color[1] = 0;
    o->sampling_mask_color[2] = 255;
}

int grrt_host_scene_from_toml_string(const char* toml_text, const char* resource_root, const GrrtCliOpts* opts,
                                     GrrtHostScene** out, char* err, size_t err_len) {
    if (!toml_text || !opts || !out) return GRRT_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    GrrtHostScene* hs = new GrrtHostScene();
    std::string msg;
    int rc = build_scene_from_toml(toml_text, resource_root, *opts, hs->data, msg);
    if (rc) {
        set_err(err, err_len, msg);
        delete hs;
        return rc;
    }
    *out = hs;
    return GRRT_OK;
}

int grrt_host_scene_from_toml_file(const char* toml_path, const char* resource_root, const GrrtCliOpts* opts,
                                   GrrtHostScene** out, char* err, size_t err_len) {
    if (!toml_path) return GRRT_ERR_INVALID_ARGUMENT;
    std::ifstream f(toml_path, std::ios::binary);
    if (!f) {  // main.rs:67-73
        set_err(err, err_len, std::string("Config file not found: ") + toml_path);
        return GRRT_ERR_INVALID_CONFIGURATION;
    }
    std::stringstream ss;
    ss << f.rdbuf();
    return grrt_host_scene_from_toml_string(ss.str().c_str(), resource_root, opts, out, err, err_len);
}

const GrrtScene* grrt_host_scene_get(const GrrtHostScene* hs) { return hs ? &hs->data.scene : nullptr; }
void grrt_host_scene_free(GrrtHostScene* hs) { delete hs; }

int grrt_host_scene_set_camera(GrrtHostScene* hs, const GrrtCliOpts* opts, char* err, size_t err_len) {
    if (!hs || !opts) return GRRT_ERR_INVALID_ARGUMENT;
    std::string msg;
    int rc = update_camera(hs->data, *opts, msg);
    if (rc) set_err(err, err_len, msg);
    return rc;
}

void grrt_host_blackbody_xyz(double temperature, double redshift, double xyz_out[3]) {
    if (xyz_out) blackbody_xyz_redshifted(temperature, redshift, xyz_out);
}
void grrt_host_xyz_to_srgb8(const double xyz[3], double exposure, uint8_t rgb_out[3]) {
    if (xyz && rgb_out) xyz_to_srgb8(xyz, exposure, rgb_out);
}

int grrt_host_ray_at(const GrrtGeometry* geometry, const double position_cartesian[3], const double direction_cartesian[3],
                     double pos_out[4], double mom_out[4], char* err, size_t err_len) {
    if (!geometry || !position_cartesian || !direction_cartesian || !pos_out || !mom_out) return GRRT_ERR_INVALID_ARGUMENT;
    std::string msg;
    int rc = ray_at(*geometry, position_cartesian, direction_cartesian, pos_out, mom_out, msg);
    if (rc) set_err(err, err_len, msg);
    return rc;
}

// IntegratedRay::save (ray.rs:35-54).  Rust's `{}` prints the shortest decimal that round-trips, without
// an exponent; %.17g round-trips too but is not byte-identical -- values compare equal, text does not.
int grrt_host_write_ray_csv(const GrrtGeometry* geometry, const GrrtStep* steps, uint64_t n, const char* path) {
    if (!geometry || (!steps && n) || !path) return GRRT_ERR_INVALID_ARGUMENT;
    FILE* f = std::fopen(path, "w");
    if (!f) return GRRT_ERR_IO;
    std::fputs("i,t,tau,x,y,z\n", f);
    for (uint64_t i = 0; i < n; ++i) {
        double c[4];
        step_to_cartesian(*geometry, steps[i].x, c);
        std::fprintf(f, "%llu,%.17g,%.17g,%.17g,%.17g,%.17g\n", (unsigned long long)steps[i].step, steps[i].t, c[0], c[1], c[2], c[3]);
    }
    return std::fclose(f) == 0 ? GRRT_OK : GRRT_ERR_IO;
}

int grrt_host_write_png_rgb8(const char* path, const uint8_t* rgb, uint32_t width, uint32_t height) {
    if (!path || !rgb) return GRRT_ERR_INVALID_ARGUMENT;
    std::string err;
    return encode_png_rgb8(path, rgb, width, height, err) ? GRRT_OK : GRRT_ERR_IO;
}
int grrt_host_write_hdr_xyz(const char* path, const double* xyza, uint32_t width, uint32_t height) {
    if (!path || !xyza) return GRRT_ERR_INVALID_ARGUMENT;
    std::string err;
    return encode_hdr_xyz(path, xyza, width, height, err) ? GRRT_OK : GRRT_ERR_IO;
}
int grrt_host_read_png_rgba8(const char* path, uint8_t** rgba_out, uint32_t* width, uint32_t* height) {
    if (!path || !rgba_out || !width || !height) return GRRT_ERR_INVALID_ARGUMENT;
    std::vector<uint8_t> px;
    std::string err;
    if (!decode_png_rgba8(path, px, *width, *height, err)) return GRRT_ERR_TEXTURE;
    uint8_t* p = (uint8_t*)std::malloc(px.size() ? px.size() : 1);
    if (!p) return GRRT_ERR_OUT_OF_MEMORY;
    std::memcpy(p, px.data(), px.size());
    *rgba_out = p;
    return GRRT_OK;
}
void grrt_host_free(void* p) { std::free(p); }

}  // extern "C"

// sizeof() of every ABI struct, so that bindings in other languages can verify their layout
// (tests/test_abi.py does it for the ctypes mirror in gr_raytracer_b200/abi.py).
extern "C" size_t grrt_abi_sizeof(int which) {
    switch (which) {
        case 0: return sizeof(GrrtGeometry);
        case 1: return sizeof(GrrtCamera);
        case 2: return sizeof(GrrtIntegration);
        case 3: return sizeof(GrrtTexture);
        case 4: return sizeof(GrrtTemperature);
        case 5: return sizeof(GrrtObject);
        case 6: return sizeof(GrrtAdaptive);
        case 7: return sizeof(GrrtScene);
        case 8: return sizeof(GrrtStep);
        case 9: return sizeof(GrrtStats);
        case 10: return sizeof(GrrtCliOpts);
        default: return 0;
    }
}

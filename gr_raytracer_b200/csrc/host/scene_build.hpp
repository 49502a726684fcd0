// scene_build.hpp -- host-side scene assembly (see scene_build.cpp)
#pragma once

#include <array>
#include <cstdint>
#include <list>
#include <map>
#include <string>
#include <vector>

#include "../../../include/grrt.h"
#include "toml.hpp"

namespace grrt_host {

struct HostBitmap {
    std::vector<uint8_t> rgba;
    uint32_t width = 0, height = 0;
};
struct HostTempLut {
    std::vector<double> r, t;
    double r_isco = 0.0;
};

// Owns every buffer the GrrtScene points into.
struct HostSceneData {
    GrrtScene scene;
    std::vector<HostBitmap> bitmaps;
    std::list<HostTempLut> temp_luts;
    std::vector<double> bb_log_t, bb_xyz;
    std::string camera_velocity_mode = "StaticObserver";  // [camera_velocity] of the scene file, kept for update_camera()
    double camera_velocity_explicit[4] = {0.0, 0.0, 0.0, 0.0};
};

int build_scene_from_toml(const std::string& text, const char* resource_root, const GrrtCliOpts& opts, HostSceneData& out,
                          std::string& err);
int update_camera(HostSceneData& scene, const GrrtCliOpts& opts, std::string& err);
int ray_at(const GrrtGeometry& g, const double position_cartesian[3], const double direction[3], double pos_out[4],
           double mom_out[4], std::string& err);
void step_to_cartesian(const GrrtGeometry& g, const double x[4], double out[4]);
void blackbody_xyz_redshifted(double temperature, double redshift, double out[3]);
void xyz_to_srgb8(const double xyz[3], double exposure, uint8_t rgb[3]);

}  // namespace grrt_host

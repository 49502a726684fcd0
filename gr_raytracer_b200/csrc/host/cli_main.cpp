// cli_main.cpp -- `grrt`: the reference's command line (src/cli/cli.rs, src/main.rs:29-178) on top of libgrrt.
//
//   grrt [--width W --height H --step-size S --max-steps N --max-radius R --epsilon E --camera-position x,y,z
//         --phi P --theta T --psi S --tone-mapping reinhard|global-linear --show-sampling-mask
//         --sampling-mask-color R,G,B] --config-file scene.toml
//        render [--filename render.png] [--from-row .. --from-col .. --to-row .. --to-col ..]
//      | render-ray --row R --col C [--filename rendered-ray.csv]
//      | render-ray-at --position x,y,z --direction x,y,z [--filename rendered-ray-at.csv]
//      | render-frames --camera-positions positions.csv [--output-pattern frame_%05d.png] [--first K --stride S --count N]
//      | blackbody --temperature T [--redshift g]
//      | blackbody-spectrum [--min-temperature 1000 --max-temperature 10000 --min-redshift 0.5 --max-redshift 2
//                            --width 1000 --height 1000 --filename blackbody_spectrum.png]
//
// Extra (not in the reference): --resource-root DIR (prefix for relative texture paths), --device N, --devices LIST
// (e.g. 0-7: one frame sharded over several GPUs inside grrt_render, grrt_init_multi), and
// render-frames: the reference's camera fly-through (scripts/Readme.md) runs the binary once per camera position;
// here one process keeps the scene resident on the GPU, replaces only the camera per frame and encodes frame k on a
// host thread while the GPU renders frame k+1.  positions.csv: one `x,y,z[,phi,theta,psi]` per line (other lines
// are skipped); frames first, first+stride, ... are rendered, so N processes with --first r --stride N --device r
// split a movie over N GPUs.
#include <chrono>
#include <fstream>
#include <algorithm>
#include <thread>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../../include/grrt.h"

namespace {

[[noreturn]] void die(const std::string& msg) {
    std::fprintf(stderr, "Error: %s\n", msg.c_str());
    std::exit(1);
}
std::vector<double> parse_list(const std::string& s, const char* what) {
    std::vector<double> out;
    size_t p = 0;
    while (p <= s.size()) {
        size_t q = s.find(',', p);
        if (q == std::string::npos) q = s.size();
        std::string tok = s.substr(p, q - p);
        char* end = nullptr;
        double v = std::strtod(tok.c_str(), &end);
        if (tok.empty() || !end || *end != '\0') die(std::string("invalid value '") + s + "' for " + what);
        out.push_back(v);
        p = q + 1;
    }
    return out;
}
bool ends_with(const std::string& s, const char* suf) {
    size_t n = std::strlen(suf);
    return s.size() >= n && s.compare(s.size() - n, n, suf) == 0;
}

// --devices: comma-separated ordinals and ranges, e.g. 0-7 or 0,2,5-6
std::vector<int> parse_devices(const std::string& spec) {
    std::vector<int> out;
    size_t p = 0;
    while (p <= spec.size()) {
        size_t q = spec.find(',', p);
        if (q == std::string::npos) q = spec.size();
        const std::string tok = spec.substr(p, q - p);
        const size_t dash = tok.find('-');
        char* end = nullptr;
        long a = std::strtol(tok.c_str(), &end, 10), b = a;
        bool ok = !tok.empty() && end != tok.c_str();
        if (ok && dash != std::string::npos) {
            ok = end == tok.c_str() + dash;
            const char* rest = tok.c_str() + dash + 1;
            b = std::strtol(rest, &end, 10);
            ok = ok && end != rest && *end == '\0' && b >= a;
        } else if (ok) {
            ok = *end == '\0';
        }
        if (!ok || a < 0 || b > 1023) die("invalid value '" + spec + "' for '--devices' (expected e.g. 0-7 or 0,2,4)");
        for (long d = a; d <= b; ++d) out.push_back((int)d);
        p = q + 1;
    }
    return out;
}

// --output-pattern holds exactly one conversion of the form %d or %0Nd (N = 1..9), which is replaced by the frame number;
// "%%" is a literal per cent sign.  The pattern never reaches printf.
std::string frame_file_name(const std::string& pattern, long long frame) {
    std::string out;
    int conversions = 0;
    for (size_t i = 0; i < pattern.size(); ++i) {
        if (pattern[i] != '%') { out += pattern[i]; continue; }
        if (i + 1 < pattern.size() && pattern[i + 1] == '%') { out += '%'; ++i; continue; }
        size_t j = i + 1;
        int width = 0;
        bool zero = false;
        if (j < pattern.size() && pattern[j] == '0') { zero = true; ++j; }
        if (j < pattern.size() && pattern[j] >= '1' && pattern[j] <= '9') { width = pattern[j] - '0'; ++j; }
        if (j >= pattern.size() || pattern[j] != 'd' || (zero && width == 0) || ++conversions > 1)
            die("invalid value '" + pattern + "' for '--output-pattern': expected one %d or %0Nd conversion");
        std::string num = std::to_string(frame);
        if ((int)num.size() < width) num = std::string((size_t)width - num.size(), zero ? '0' : ' ') + num;
        out += num;
        i = j;
    }
    if (conversions != 1) die("invalid value '" + pattern + "' for '--output-pattern': expected one %d or %0Nd conversion");
    return out;
}
}  // namespace

int main(int argc, char** argv) {
    GrrtCliOpts opts;
    grrt_cli_opts_default(&opts);
    std::string config_file, action, filename, resource_root, positions_file, output_pattern = "frame_%05d.png";
    long long first_frame = 0, frame_stride = 1, frame_count = -1;
    double bb_temperature = -1.0, bb_redshift = 1.0, bb_tmin = 1000.0, bb_tmax = 10000.0, bb_gmin = 0.5, bb_gmax = 2.0;
    bool size_given = false, short_r_given = false;
    std::string short_r_value;
    long long from_row = -1, from_col = -1, to_row = -1, to_col = -1, ray_row = -1, ray_col = -1;
    std::vector<double> at_pos, at_dir;
    int device = 0;
    std::vector<int> devices;   // --devices 0-7 | 0,2,4: the frame is sharded over these GPUs inside grrt_render

    std::vector<std::string> args(argv + 1, argv + argc);
    auto value = [&](size_t& i, const std::string& flag) -> std::string {
        // --flag=value or --flag value
        const std::string& a = args[i];
        size_t eq = a.find('=');
        if (eq != std::string::npos) return a.substr(eq + 1);
        if (i + 1 >= args.size()) die("a value is required for '" + flag + "'");
        return args[++i];
    };
    for (size_t i = 0; i < args.size(); ++i) {
        std::string a = args[i];
        std::string flag = a.substr(0, a.find('='));
        if (flag == "--width") { opts.width = std::atoll(value(i, flag).c_str()); size_given = true; }
        else if (flag == "--height") { opts.height = std::atoll(value(i, flag).c_str()); size_given = true; }
        else if (flag == "--temperature" || flag == "-t") bb_temperature = std::atof(value(i, flag).c_str());
        else if (flag == "--redshift") bb_redshift = std::atof(value(i, flag).c_str());
        else if (flag == "--min-temperature") bb_tmin = std::atof(value(i, flag).c_str());
        else if (flag == "--max-temperature") bb_tmax = std::atof(value(i, flag).c_str());
        else if (flag == "--min-redshift") bb_gmin = std::atof(value(i, flag).c_str());
        else if (flag == "--max-redshift") bb_gmax = std::atof(value(i, flag).c_str());
        else if (flag == "--step-size") opts.step_size = std::atof(value(i, flag).c_str());
        else if (flag == "--max-steps") opts.max_steps = std::strtoull(value(i, flag).c_str(), nullptr, 10);
        else if (flag == "--max-radius") opts.max_radius = std::atof(value(i, flag).c_str());
        else if (flag == "--epsilon") opts.epsilon = std::atof(value(i, flag).c_str());
        else if (flag == "--camera-position") {
            std::vector<double> v = parse_list(value(i, flag), "--camera-position");
            if (v.size() != 3) die("Invalid configuration: Camera position must be a vector of length 3");
            for (int k = 0; k < 3; ++k) opts.camera_position[k] = v[k];
        } else if (flag == "--phi") opts.phi = std::atof(value(i, flag).c_str());
        else if (flag == "--theta") opts.theta = std::atof(value(i, flag).c_str());
        else if (flag == "--psi") opts.psi = std::atof(value(i, flag).c_str());
        else if (flag == "--tone-mapping") {
            std::string v = value(i, flag);
            if (v == "reinhard") opts.tone_mapping = GRRT_TONEMAP_REINHARD;
            else if (v == "global-linear") opts.tone_mapping = GRRT_TONEMAP_GLOBAL_LINEAR;
            else die("invalid value '" + v + "' for '--tone-mapping' [possible values: reinhard, global-linear]");
        } else if (flag == "--show-sampling-mask") opts.show_sampling_mask = 1;
        else if (flag == "--sampling-mask-color") {
            std::string raw = value(i, flag);
            std::vector<double> v = parse_list(raw, "--sampling-mask-color");
            if (v.size() != 3) die("invalid RGB color '" + raw + "'; expected R,G,B (for example 255,0,255)");
            for (int k = 0; k < 3; ++k) {
                if (v[k] < 0 || v[k] > 255 || v[k] != (double)(int)v[k]) die("invalid RGB color '" + raw + "'; expected three values from 0 to 255");
                opts.sampling_mask_color[k] = (uint8_t)v[k];
            }
        } else if (flag == "--config-file" || flag == "-c") config_file = value(i, flag);
        else if (flag == "--resource-root") resource_root = value(i, flag);
        else if (flag == "--device") device = std::atoi(value(i, flag).c_str());
        else if (flag == "--devices") devices = parse_devices(value(i, flag));
        else if (flag == "--filename") filename = value(i, flag);
        else if (flag == "--from-row") from_row = std::atoll(value(i, flag).c_str());
        else if (flag == "--from-col") from_col = std::atoll(value(i, flag).c_str());
        else if (flag == "--to-row") to_row = std::atoll(value(i, flag).c_str());
        else if (flag == "--to-col") to_col = std::atoll(value(i, flag).c_str());
        else if (flag == "--row") ray_row = std::atoll(value(i, flag).c_str());
        else if (flag == "-r") { short_r_value = value(i, flag); short_r_given = true; }
        else if (flag == "--col") ray_col = std::atoll(value(i, flag).c_str());
        else if (flag == "--position" || flag == "-p") at_pos = parse_list(value(i, flag), "--position");
        else if (flag == "--direction" || flag == "-d") at_dir = parse_list(value(i, flag), "--direction");
        else if (flag == "--camera-positions") positions_file = value(i, flag);
        else if (flag == "--output-pattern") { output_pattern = value(i, flag); frame_file_name(output_pattern, 0); }  // validates
        else if (flag == "--first") first_frame = std::atoll(value(i, flag).c_str());
        else if (flag == "--stride") frame_stride = std::atoll(value(i, flag).c_str());
        else if (flag == "--count") frame_count = std::atoll(value(i, flag).c_str());
        else if (a == "render" || a == "render-ray" || a == "render-ray-at" || a == "render-frames" || a == "blackbody" ||
                 a == "blackbody-spectrum") action = a;
        else die("unexpected argument '" + a + "'");
    }
    if (action.empty()) die("a subcommand is required: render | render-ray | render-ray-at | render-frames | blackbody | blackbody-spectrum");
    if (short_r_given) {  // clap: -r is --row of render-ray and --redshift of blackbody
        if (action == "blackbody") bb_redshift = std::atof(short_r_value.c_str());
        else ray_row = std::atoll(short_r_value.c_str());
    }

    // ---- cli/blackbody.rs: no scene; `blackbody` needs no GPU either ----
    if (action == "blackbody") {
        if (bb_temperature < 0.0) die("the following required arguments were not provided: --temperature <TEMPERATURE>");
        double xyz[3];
        uint8_t c[3];
        grrt_host_blackbody_xyz(bb_temperature, bb_redshift, xyz);
        grrt_host_xyz_to_srgb8(xyz, 1.0, c);
        std::printf("Blackbody color at T=%gK (redshift=%g):\n", bb_temperature, bb_redshift);
        std::printf("XYZ:  %.4f, %.4f, %.4f\n", xyz[0], xyz[1], xyz[2]);
        std::printf("sRGB: R=%d, G=%d, B=%d\n", c[0], c[1], c[2]);
        std::printf("sRGB: R=%.4f, G=%.4f, B=%.4f\n", c[0] / 255.0, c[1] / 255.0, c[2] / 255.0);
        std::printf("Color block: \x1b[48;2;%d;%d;%dm      \x1b[0m\n", c[0], c[1], c[2]);
        return 0;
    }
    if (action == "blackbody-spectrum") {
        const uint32_t W = size_given ? (uint32_t)opts.width : 1000u, H = size_given ? (uint32_t)opts.height : 1000u;
        if (W == 0 || H == 0) die("Invalid configuration: --width and --height must be positive");
        if (filename.empty()) filename = "blackbody_spectrum.png";
        const size_t n = (size_t)W * H;
        std::vector<double> xyza(4 * n);
        unsigned nt = std::max(1u, std::thread::hardware_concurrency());
        std::vector<std::thread> pool;
        for (unsigned t = 0; t < nt; ++t)
            pool.emplace_back([&, t] {
                for (size_t i = t; i < n; i += nt) {  // temperature along x, redshift along y (blackbody.rs:62-71)
                    double x = (double)(i % W), y = (double)(i / W);
                    double temperature = bb_tmin + x * (bb_tmax - bb_tmin) / ((double)W - 1.0);
                    double redshift = bb_gmin + y * (bb_gmax - bb_gmin) / ((double)H - 1.0);
                    grrt_host_blackbody_xyz(temperature, redshift, &xyza[4 * i]);
                    xyza[4 * i + 3] = 1.0;
                }
            });
        for (auto& th : pool) th.join();
        GrrtCtx* bctx = nullptr;
        if (grrt_init(device, &bctx) != 0) die("no usable CUDA device (sm_100): grrt has no CPU fallback");
        std::vector<uint8_t> rgb(3 * n);
        if (grrt_tonemap(bctx, xyza.data(), n, opts.tone_mapping, 1.0, rgb.data())) die(grrt_last_error(bctx));
        grrt_destroy(bctx);
        if (grrt_host_write_png_rgb8(filename.c_str(), rgb.data(), W, H)) die("I/O error: cannot write " + filename);
        std::printf("Saved blackbody spectrum to %s\n", filename.c_str());
        return 0;
    }
    if (config_file.empty()) die("Invalid configuration: Config file is required for this action");

    auto t0 = std::chrono::steady_clock::now();
    char err[1024] = {0};
    GrrtHostScene* hs = nullptr;
    int rc = grrt_host_scene_from_toml_file(config_file.c_str(), resource_root.empty() ? nullptr : resource_root.c_str(), &opts, &hs,
                                            err, sizeof err);
    if (rc) die(err);
    const GrrtScene* scene = grrt_host_scene_get(hs);
    GrrtCtx* ctx = nullptr;
    if (!devices.empty()) {
        if (grrt_init_multi((int)devices.size(), devices.data(), &ctx) != 0)
            die("--devices: needs 1.." + std::to_string(GRRT_MAX_PEERS) + " distinct sm_100 CUDA devices with peer access to the first");
    } else if (grrt_init(device, &ctx) != 0) {
        die("no usable CUDA device (sm_100): grrt has no CPU fallback");
    }

    if (action == "render-frames") {
        if (positions_file.empty()) die("render-frames needs --camera-positions");
        if (frame_stride <= 0 || first_frame < 0) die("Invalid configuration: --first must be >= 0 and --stride > 0");
        std::vector<std::vector<double>> cams;  // x,y,z[,phi,theta,psi]
        {
            std::ifstream in(positions_file);
            if (!in) die("I/O error: cannot read " + positions_file);
            std::string line;
            while (std::getline(in, line)) {
                std::vector<double> v;
                size_t p = 0;
                bool ok = !line.empty();
                while (ok && p <= line.size()) {
                    size_t q = line.find(',', p);
                    if (q == std::string::npos) q = line.size();
                    std::string tok = line.substr(p, q - p);
                    char* end = nullptr;
                    double x = std::strtod(tok.c_str(), &end);
                    while (end && (*end == ' ' || *end == '\r' || *end == '\t')) ++end;
                    if (tok.empty() || !end || *end != '\0') ok = false;
                    else v.push_back(x);
                    p = q + 1;
                }
                if (ok && (v.size() == 3 || v.size() == 6)) cams.push_back(v);  // header / comment lines are skipped
            }
        }
        if (cams.empty()) die("Invalid configuration: no camera positions in " + positions_file);
        GrrtDeviceScene* ds = nullptr;
        if (grrt_scene_upload(ctx, scene, &ds)) die(grrt_last_error(ctx));
        const uint32_t W = (uint32_t)scene->camera.cols, H = (uint32_t)scene->camera.rows;
        const size_t n = (size_t)W * H;
        std::vector<double> xyza(4 * n);
        std::vector<uint8_t> rgb[2] = {std::vector<uint8_t>(3 * n), std::vector<uint8_t>(3 * n)};
        std::thread writer;
        std::string write_error;
        long long done = 0;
        double device_ms = 0.0;
        for (long long k = first_frame; k < (long long)cams.size() && (frame_count < 0 || done < frame_count); k += frame_stride, ++done) {
            const std::vector<double>& c = cams[(size_t)k];
            for (int j = 0; j < 3; ++j) opts.camera_position[j] = c[j];
            if (c.size() == 6) { opts.phi = c[3]; opts.theta = c[4]; opts.psi = c[5]; }
            if (grrt_host_scene_set_camera(hs, &opts, err, sizeof err)) die(std::string("frame ") + std::to_string(k) + ": " + err);
            if (grrt_scene_set_camera(ctx, ds, &scene->camera)) die(grrt_last_error(ctx));
            GrrtStats st;
            if (grrt_render_resident(ctx, ds, 0, 0, H, W, xyza.data(), nullptr, nullptr, nullptr, &st)) die(grrt_last_error(ctx));
            device_ms += st.device_ms_total;
            std::vector<uint8_t>& buf = rgb[done & 1];
            if (grrt_tonemap(ctx, xyza.data(), n, opts.tone_mapping, 1.0, buf.data())) die(grrt_last_error(ctx));
            const std::string name = frame_file_name(output_pattern, k);
            if (writer.joinable()) writer.join();  // frame k-1 was encoding while this one rendered
            if (!write_error.empty()) die(write_error);
            std::string path = name;
            writer = std::thread([&buf, &write_error, path, W, H] {
                if (grrt_host_write_png_rgb8(path.c_str(), buf.data(), W, H)) write_error = "I/O error: cannot write " + path;
            });
        }
        if (writer.joinable()) writer.join();
        if (!write_error.empty()) die(write_error);
        grrt_scene_free(ctx, ds);
        std::fprintf(stderr, "[grrt] rendered %lld frames (%ux%u), device %.1f ms per frame\n", done, W, H, done ? device_ms / done : 0.0);
    } else if (action == "render") {
        if (filename.empty()) filename = "render.png";
        uint32_t fr = from_row < 0 ? 0 : (uint32_t)from_row, fc = from_col < 0 ? 0 : (uint32_t)from_col;
        uint32_t tr = to_row < 0 ? (uint32_t)scene->camera.rows : (uint32_t)to_row;
        uint32_t tc = to_col < 0 ? (uint32_t)scene->camera.cols : (uint32_t)to_col;
        if (tr < fr || tc < fc) die("Invalid configuration: empty section");
        size_t n = (size_t)(tr - fr) * (tc - fc);
        std::vector<double> xyza(4 * n);
        GrrtStats st;
        rc = grrt_render(ctx, scene, fr, fc, tr, tc, xyza.data(), nullptr, nullptr, nullptr, &st);
        if (rc) die(grrt_last_error(ctx));
        if (ends_with(filename, ".hdr")) {  // raytracer.rs:468-480
            if (grrt_host_write_hdr_xyz(filename.c_str(), xyza.data(), tc - fc, tr - fr)) die("I/O error: cannot write " + filename);
        } else {
            std::vector<uint8_t> rgb(3 * n);
            rc = grrt_tonemap(ctx, xyza.data(), n, opts.tone_mapping, 1.0, rgb.data());
            if (rc) die(grrt_last_error(ctx));
            if (grrt_host_write_png_rgb8(filename.c_str(), rgb.data(), tc - fc, tr - fr)) die("I/O error: cannot write " + filename);
        }
        std::fprintf(stderr,
                     "[grrt] %llu base rays + %llu sub-sample rays (%llu pixels supersampled), %llu RKF45 steps, %llu attempts, "
                     "%llu ray-march samples, %llu ray errors; device %.1f ms\n",
                     (unsigned long long)st.rays_base, (unsigned long long)st.rays_supersample, (unsigned long long)st.pixels_supersampled,
                     (unsigned long long)st.accepted_steps, (unsigned long long)st.attempts, (unsigned long long)st.raymarch_samples,
                     (unsigned long long)st.ray_errors, st.device_ms_total);
        std::fprintf(stderr, "[grrt] saved image to %s\n", filename.c_str());
    } else {
        std::vector<GrrtStep> steps(scene->integration.max_steps ? scene->integration.max_steps : 1);
        uint64_t n = 0;
        int32_t stop = 0;
        if (action == "render-ray") {
            if (ray_row < 0 || ray_col < 0) die("render-ray needs --row and --col");
            if (filename.empty()) filename = "rendered-ray.csv";
            rc = grrt_trace_pixel_ray(ctx, scene, ray_row, ray_col, steps.data(), steps.size(), &n, &stop);
        } else {
            if (at_pos.size() != 3) die("Invalid configuration: Position must be a vector of length 3");
            if (at_dir.size() != 3) die("Invalid configuration: Direction must be a vector of length 3");
            if (filename.empty()) filename = "rendered-ray-at.csv";
            double pos[4], mom[4];
            if (grrt_host_ray_at(&scene->geometry, at_pos.data(), at_dir.data(), pos, mom, err, sizeof err)) die(err);
            rc = grrt_trace_ray(ctx, scene, pos, mom, steps.data(), steps.size(), &n, &stop);
        }
        if (rc < 0) die(grrt_last_error(ctx));
        if (rc > 0) die("Integration error: Max steps reached");  // IntegrationError::MaxStepsReached
        if (n > steps.size()) n = steps.size();
        if (grrt_host_write_ray_csv(&scene->geometry, steps.data(), n, filename.c_str())) die("I/O error: cannot write " + filename);
        static const char* names[] = {"None", "HorizonReached", "CelestialSphereReached", "CoordinateIsNan", "ClosedOrbitDetected", "Error"};
        std::fprintf(stderr, "[grrt] Stop reason: %s; saved integrated ray (%llu steps) to %s\n", names[stop < 0 || stop > 5 ? 5 : stop],
                     (unsigned long long)n, filename.c_str());
    }
    grrt_destroy(ctx);
    grrt_host_scene_free(hs);
    double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::fprintf(stderr, "[grrt] Elapsed time: %.2fs\n", secs);
    return 0;
}

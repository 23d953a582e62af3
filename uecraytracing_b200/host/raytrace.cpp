// raytrace.cpp -- the drop-in binary: same CLI contract as the reference's main()
// (source.cpp:190-231): raytrace [-h] [-v] [-l N] [-o FILE | FILE]; no output -> help, exit 0;
// stdout lines "rendering...", "rendering finished", "write to file : <f>", "success";
// a failed write prints "error" and exits 1; an unknown option aborts (uncaught exception).
// New, additive flags (north_star): --width --spp --depth --seed, plus --gpus --scene --dense-n
// --exact-hits --scan --dump-accum --stats, and --frames N --orbit DEG: N frames of a camera orbit through the frame queue
// (K1 of frame k + 1 overlaps the D2H and the PNG write of frame k; files <output stem>_0000.png ...).
// Defaults equal the reference's macros (400 / 100 / 50).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <random>
#include <chrono>
#include <string>
#include <thread>
#include <vector>

#include "cli_options.hpp"
#include "png_writer.hpp"
#include "yk/render.hpp"
#include "yk/scene_io.hpp"
#include "yk/scenes.hpp"

int main(int argc, char* argv[]) {
  std::ios::sync_with_stdio(false);
  std::cout.tie(nullptr);

  uecrt_host::options options("raytrace", "raytracing program");
  options.add('h', "help", "print usage", false)
      .add('v', "verbose", "verbose output", false)
      .add('o', "output", "filename of output", true)
      .add('l', "verbose-level", "set verbose level", true)
      .add(0, "stats", "print the kernel timing line (K1 / reduce / K3 ms, samples/s) without the per-pixel lines of -v", false)
      .add(0, "width", "image width (default 400; height = width*9/16)", true)
      .add(0, "spp", "samples per pixel (default 100)", true)
      .add(0, "depth", "maximum bounce depth (default 50)", true)
      .add(0, "seed", "Philox seed (default: random)", true)
      .add(0, "gpus", "number of B200s to shard samples over (default 1)", true)
      .add(0, "scene", "ref4 | ref4d | final | dense (default ref4)", true)
      .add(0, "scene-file", "render the spheres (and camera) described in a text file (see host/yk/scene_io.hpp)", true)
      .add(0, "scene-seed", "seed of the random scene generator (default 1)", true)
      .add(0, "dense-n", "sphere count of --scene dense (default 10000)", true)
      .add(0, "exact-hits", "resolve every closest hit in FP64", false)
      .add(0, "scan", "closest-hit search: auto | fp32 | tensor | grid (default auto: brute force on the tensor cores from 96 spheres up; grid = opt-in spatial index)", true)
      .add(0, "frames", "render N frames of a camera orbit about the y axis, PNG write of frame k overlapped with frame k + 1 (default 1)", true)
      .add(0, "orbit", "degrees the camera turns per frame of --frames (default 2)", true)
      .add(0, "dump-accum", "also write the linear float image (.pfm) or the raw W*H*4 f32 accumulation buffer", true);
  options.positional("output");
  options.parse(argc, argv);
  if (options.count("help") || !options.count("output")) {
    std::cout << options.help() << std::endl;
    std::exit(EXIT_SUCCESS);
  }
  if (options.count("verbose") && !yk::verbose) ++yk::verbose;
  if (options.count("verbose-level")) yk::verbose = options.as<std::uint32_t>("verbose-level");
  yk::stats = options.count("stats") != 0;

  const std::string filename = options.last("output");

  yk::render_config cfg;
  if (options.count("width")) cfg.width = options.as<std::uint32_t>("width");
  if (options.count("spp")) cfg.spp = options.as<std::uint32_t>("spp");
  if (options.count("depth")) cfg.max_depth = options.as<std::uint32_t>("depth");
  cfg.seed = options.count("seed") ? options.as<std::uint64_t>("seed")
                                   : (std::uint64_t(std::random_device{}()) << 32) ^ std::random_device{}();   // source.cpp:159
  if (options.count("gpus")) cfg.gpus = options.as<int>("gpus");
  cfg.exact_hits = options.count("exact-hits") != 0;
  if (options.count("scan")) {
    const std::string m = options.last("scan");
    if (m == "fp32") cfg.scan = UECRT_SCAN_FP32;
    else if (m == "tensor") cfg.scan = UECRT_SCAN_TENSOR;
    else if (m == "grid") cfg.scan = UECRT_SCAN_GRID;
    else if (m != "auto") { std::cout << "error" << std::endl; std::cerr << "--scan must be auto, fp32, tensor or grid" << std::endl; std::exit(EXIT_FAILURE); }
  }
  const std::string scene = options.count("scene") ? options.last("scene") : "ref4";
  const std::uint32_t scene_seed = options.count("scene-seed") ? options.as<std::uint32_t>("scene-seed") : 1u;
  const std::uint32_t dense_n = options.count("dense-n") ? options.as<std::uint32_t>("dense-n") : 10000u;

  yk::hittable_list<double> world;
  yk::camera<double> cam;
  if (options.count("scene-file")) {
    try {
      yk::scene_io::scene_file f = yk::scene_io::load(options.last("scene-file"));
      world = f.world;
      cam = f.cam;
    } catch (const std::exception& e) {
      std::cout << "error" << std::endl;
      std::cerr << e.what() << std::endl;
      std::exit(EXIT_FAILURE);
    }
  } else if (scene == "ref4") { world = yk::scenes::reference_scene(); cam = yk::scenes::reference_camera(); }
  else if (scene == "ref4d") { world = yk::scenes::reference_scene_dielectric(); cam = yk::scenes::reference_camera(); }
  else if (scene == "final") { world = yk::scenes::final_scene(scene_seed); cam = yk::scenes::final_camera(); }
  else if (scene == "dense") { world = yk::scenes::dense_scene(dense_n, scene_seed); cam = yk::scenes::dense_camera(); }
  else { std::cout << "unknown scene : " << scene << std::endl; std::exit(EXIT_FAILURE); }

  const std::uint32_t frames = options.count("frames") ? options.as<std::uint32_t>("frames") : 1u;
  if (frames > 1) {
    // a sequence: <stem>_0000.png ... ; the PNG of frame k is deflated and written by a host thread while K1 of frame k + 1 runs
    const double step = options.count("orbit") ? std::stod(options.last("orbit")) : 2.0;
    std::vector<yk::camera<double>> cams;
    for (std::uint32_t k = 0; k < frames; ++k) cams.push_back(yk::orbit_y(cam, step * k));
    const std::size_t dot = filename.rfind('.');
    const std::string stem = dot == std::string::npos ? filename : filename.substr(0, dot), ext = dot == std::string::npos ? ".png" : filename.substr(dot);
    const std::uint32_t W = cfg.width, H = yk::constants::height_for(cfg.width);
    std::vector<std::vector<std::uint8_t>> copies(2, std::vector<std::uint8_t>((std::size_t)W * H * 3));
    std::thread writer[2];
    bool write_ok = true;
    double k1_sum = 0.0;
    std::cout << "rendering..." << std::endl;
    double wall = 0.0;
    try {
      wall = yk::render_sequence(world, cams, cfg, [&](std::size_t k, const std::uint8_t* rgb, float k1_ms) {
        std::thread& t = writer[k & 1];
        if (t.joinable()) t.join();                       // the write of frame k - 2
        std::copy(rgb, rgb + copies[k & 1].size(), copies[k & 1].begin());
        k1_sum += k1_ms;
        char num[16];
        std::snprintf(num, sizeof num, "_%04zu", k);
        const std::string path = stem + num + ext;
        t = std::thread([&, path, k] {
          if (!uecrt_host::write_png(path.c_str(), (int)W, (int)H, 3, copies[k & 1].data(), (int)(3 * W))) write_ok = false;
        });
        if (yk::verbose || yk::stats) std::cout << "frame " << k << ": K1 " << k1_ms << " ms -> " << path << std::endl;
      });
    } catch (const std::exception& e) {
      for (auto& t : writer) if (t.joinable()) t.join();
      std::cout << "error" << std::endl;
      std::cerr << e.what() << std::endl;
      std::exit(EXIT_FAILURE);
    }
    const auto tj = std::chrono::steady_clock::now();
    for (auto& t : writer) if (t.joinable()) t.join();
    wall += std::chrono::duration<double>(std::chrono::steady_clock::now() - tj).count();
    std::cout << "rendering finished" << std::endl;
    if (yk::verbose || yk::stats)
      std::cout << frames << " frames in " << wall * 1e3 << " ms wall-clock incl. PNG (" << wall * 1e3 / frames << " ms per frame; K1 alone " << k1_sum / frames
                << " ms per frame)" << std::endl;
    std::cout << "write to file : " << stem << "_0000" << ext << " .. " << frames << " frames" << std::endl;
    if (!write_ok) { std::cout << "error" << std::endl; std::exit(EXIT_FAILURE); }
    std::cout << "success" << std::endl;
    return 0;
  }

  // rendering
  const bool dump = options.count("dump-accum") != 0;
  yk::image_t image;
  try {
    image = yk::render(world, cam, cfg, dump);
  } catch (const std::exception& e) {
    std::cout << "error" << std::endl;
    std::cerr << e.what() << std::endl;
    std::exit(EXIT_FAILURE);
  }

  std::cout << "write to file : " << filename << std::endl;
  if (!uecrt_host::write_png(filename.c_str(), (int)image.width, (int)image.height, 3, image.data(), (int)(3 * image.width))) {
    std::cout << "error" << std::endl;
    std::exit(EXIT_FAILURE);
  }
  if (dump) {
    const std::string path = options.last("dump-accum");
    bool ok;
    if (path.size() > 4 && path.substr(path.size() - 4) == ".pfm") {
      ok = uecrt_host::write_pfm(path.c_str(), (int)image.width, (int)image.height, image.accum.data(), cfg.spp) != 0;
    } else {
      std::FILE* f = std::fopen(path.c_str(), "wb");
      ok = f && std::fwrite(image.accum.data(), sizeof(float), image.accum.size(), f) == image.accum.size();
      if (f) ok = (std::fclose(f) == 0) && ok;
    }
    if (!ok) {
      std::cout << "error" << std::endl;
      std::exit(EXIT_FAILURE);
    }
  }
  std::cout << "success" << std::endl;
}

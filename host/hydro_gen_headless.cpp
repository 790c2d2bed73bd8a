// hydro_gen_headless.cpp — the erosion part of the reference's main() without the window:
// reads the reference's config.ini ([window] width,height  [map] size  [erosion] type,
// particle_count; written with the same defaults when missing, src/main.cpp:203-234), sets up
// settings / textures / heightmap / programs in the reference's order (main.cpp:252-276) and
// runs the per-step dispatch of main.cpp:310-324 for --steps iterations with an injected
// clock (--dt per iteration) and seed.  Prints the device-timed step rate and the mass sums.
//
//   hydro-gen-headless [--config config.ini] [--steps 500] [--seed 1234.5] [--dt 0.015]
//                      [--rain-period N] [--no-rain] [--device 0] [--dump prefix]
//                      [--resume file.hgck] [--checkpoint file.hgck]
// --resume loads fields, settings and the step counter from a checkpoint (hg_checkpoint_load)
// instead of generating a heightmap; --checkpoint writes one after the last step.
#include <cctype>
#include <cstring>
#include <fstream>
#include <map>
#include <string>
#include "hydrogen_erosion.hpp"

namespace {

// Minimal INI reader with inih's behaviour for the five keys the reference reads:
// [section], key = value, ';' or '#' comments (whole-line or after whitespace), names case-insensitive.
struct Ini {
    std::map<std::string, std::string> kv;
    bool ok = false;
    static std::string trim(std::string s) {
        size_t a = 0, b = s.size();
        while (a < b && std::isspace((unsigned char)s[a])) a++;
        while (b > a && std::isspace((unsigned char)s[b - 1])) b--;
        return s.substr(a, b - a);
    }
    static std::string lower(std::string s) { for (auto& c : s) c = (char)std::tolower((unsigned char)c); return s; }
    explicit Ini(const std::string& path) {
        std::ifstream f(path);
        if (!f) return;
        ok = true;
        std::string line, section;
        while (std::getline(f, line)) {
            for (size_t i = 0; i < line.size(); i++)
                if ((line[i] == ';' || line[i] == '#') && (i == 0 || std::isspace((unsigned char)line[i - 1]))) { line.resize(i); break; }
            line = trim(line);
            if (line.empty()) continue;
            if (line.front() == '[' && line.back() == ']') { section = lower(trim(line.substr(1, line.size() - 2))); continue; }
            size_t eq = line.find_first_of("=:");
            if (eq == std::string::npos) continue;
            kv[section + "=" + lower(trim(line.substr(0, eq)))] = trim(line.substr(eq + 1));
        }
    }
    std::string get(const std::string& s, const std::string& k, const std::string& def) const {
        auto it = kv.find(lower(s) + "=" + lower(k));
        return it == kv.end() ? def : it->second;
    }
    u32 get_unsigned(const std::string& s, const std::string& k, u32 def) const {
        std::string v = get(s, k, "");
        if (v.empty()) return def;
        char* end = nullptr;
        unsigned long long x = std::strtoull(v.c_str(), &end, 0);
        return end > v.c_str() ? (u32)x : def;
    }
};

const char* DEFAULT_CONFIG =          // src/main.cpp:206-216
    "[window]\nwidth = 1280\nheight = 720\n\n[map]\nsize=1024\n\n[erosion]\n"
    "; type = grid or type = particle\ntype = grid\n"
    "; particle_count works only when the erosion type is \"particle\"\nparticle_count = 262144";

}  // namespace

int main(int argc, char** argv) {
    std::string config = "config.ini", dump, resume, checkpoint;
    u32 steps = 500;
    float seed = 1234.5f, dt = 0.015f;
    int device = 0, rain_period = -1;
    bool no_rain = false;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto next = [&]() -> const char* { if (i + 1 >= argc) { std::fprintf(stderr, "%s needs a value\n", a.c_str()); std::exit(2); } return argv[++i]; };
        if (a == "--config") config = next();
        else if (a == "--steps") steps = (u32)std::strtoul(next(), nullptr, 0);
        else if (a == "--seed") seed = std::strtof(next(), nullptr);
        else if (a == "--dt") dt = std::strtof(next(), nullptr);
        else if (a == "--device") device = std::atoi(next());
        else if (a == "--rain-period") rain_period = std::atoi(next());
        else if (a == "--no-rain") no_rain = true;
        else if (a == "--dump") dump = next();
        else if (a == "--resume") resume = next();
        else if (a == "--checkpoint") checkpoint = next();
        else { std::fprintf(stderr, "unknown argument %s\n", a.c_str()); return 2; }
    }

    Ini ini(config);
    if (!ini.ok) {                                   // main.cpp:203-218
        std::fprintf(stderr, "Failed to load %s, creating a new default config file...\n", config.c_str());
        std::ofstream(config) << DEFAULT_CONFIG;
        ini = Ini(config);
        if (!ini.ok) { std::fprintf(stderr, "Failed to write %s\n", config.c_str()); return 1; }
    }
    const u32 MAP_SIZE = ini.get_unsigned("map", "size", 1024);
    Erosion::Programs::Erosion_type erosion_type = Erosion::Programs::GRID;
    u32 particle_count = 0;
    if (ini.get("erosion", "type", "grid") == "particle") {   // main.cpp:226-234
        erosion_type = Erosion::Programs::PARTICLES;
        particle_count = ini.get_unsigned("erosion", "particle_count", 262144);
    }

    State::Program_state state;
    state.should_rain = !no_rain;
    State::Settings settings = State::setup_settings(erosion_type == Erosion::Programs::PARTICLES, particle_count);
    settings.map.data.seed = seed;
    if (rain_period > 0) settings.rain.data.period = rain_period;
    Compute_program comput_map("heightmap.glsl");
    State::World::Textures world_data = State::World::gen_textures(MAP_SIZE, particle_count, device);
    State::World::gen_heightmap(settings, world_data, comput_map);
    Erosion::Programs* erosion_progs = Erosion::setup_shaders(erosion_type, settings, world_data, particle_count);
    u32 step0 = 0;
    if (!resume.empty()) {                           // fields, the three settings structs and the step counter
        hydrogen_detail::check(hg_checkpoint_load(world_data.ctx, resume.c_str()), "hg_checkpoint_load");
        hg_get_erosion(world_data.ctx, &settings.erosion.data);
        hg_get_rain(world_data.ctx, &settings.rain.data);
        hg_get_map(world_data.ctx, &settings.map.data);
        hg_get_steps(world_data.ctx, &step0);
        state.erosion_steps = step0;
    }

    hg_sync(world_data.ctx);
    hg_timer_start(world_data.ctx);
    for (u32 k = 0; k < steps; k++) {                // main.cpp:310-324
        world_data.time = (float)(step0 + k + 1) * dt;
        state.erosion_steps++;
        if (erosion_type == Erosion::Programs::GRID) {
            if (state.should_rain) {
                if (!(state.erosion_steps % (u32)settings.rain.data.period)) Erosion::dispatch_grid_rain(*erosion_progs, world_data);
            }
            Erosion::dispatch_grid(*erosion_progs, world_data);
        } else {
            Erosion::dispatch_particle(*erosion_progs, world_data, state.should_rain);
        }
    }
    float ms = 0.f;
    hg_timer_stop(world_data.ctx, &ms);
    double mass[5];
    hydrogen_detail::check(hg_mass(world_data.ctx, mass), "hg_mass");
    std::printf("%s %ux%u, %u steps: %.3f ms/step, %.3f Gcell-steps/s; kernels launched %llu\n",
                erosion_type == Erosion::Programs::GRID ? "grid" : "particle", MAP_SIZE, MAP_SIZE, steps, ms / (steps ? steps : 1),
                steps ? (double)MAP_SIZE * MAP_SIZE * steps / (ms * 1e-3) / 1e9 : 0.0, (unsigned long long)hg_launch_count(world_data.ctx));
    std::printf("mass: rock %.6f dirt %.6f water %.6f sediment %.6f %.6f\n", mass[0], mass[1], mass[2], mass[3], mass[4]);
    if (!checkpoint.empty()) {
        hg_set_steps(world_data.ctx, state.erosion_steps);
        hydrogen_detail::check(hg_checkpoint_save(world_data.ctx, checkpoint.c_str()), "hg_checkpoint_save");
    }
    if (!dump.empty()) {                             // raw little-endian RGBA32F, one file per renderer input
        const int fields[2] = {HG_FIELD_HEIGHTMAP, HG_FIELD_SEDIMENT};
        const char* names[2] = {"heightmap", "sediment"};
        for (int f = 0; f < 2; f++) {
            std::vector<float> img = State::World::download(world_data, fields[f]);
            std::ofstream o(dump + "." + names[f] + ".rgba32f", std::ios::binary);
            o.write(reinterpret_cast<const char*>(img.data()), (std::streamsize)(img.size() * sizeof(float)));
        }
    }
    delete erosion_progs;
    State::World::delete_textures(world_data);
    State::delete_settings(settings);
    return 0;
}

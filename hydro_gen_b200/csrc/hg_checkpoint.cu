// hg_checkpoint.cu — field checkpoint of a context.  The reference has no
// dump format: its state lives in GL textures and is lost at exit (src/main.cpp:334-345 only
// deletes them).  This one holds exactly what a resumed run needs to continue bit for bit:
// the three settings structs (byte images of bindings.glsl:39-99), the step counter that drives
// the rain schedule (src/main.cpp:315-319), the fields in the reference's texture format
// (RGBA32F, little endian, [y][x][4]) and, in particle mode, the droplet SSBO.
//
//   HgCkptHeader (fixed, little endian)  |  json_bytes of informational JSON  |  fields  |  particles
//
// The binary header is what hg_checkpoint_load reads; the JSON block repeats it for people and
// tools (hydro_gen_b200/checkpoint.py reads both).  A slab saves its own rows; load requires a
// context of the same geometry and mode, and on a connected slab ends with a collective halo refresh
// (hg_slab_refresh_halo), so every rank loads its file at the same point.
#include <cstddef>
#include <cstdio>
#include <string>
#include <vector>
#include "hg_internal.cuh"

namespace {

struct HgCkptHeader {
    char magic[8];               // "HGCKPT01"
    uint32_t version;            // 1
    uint32_t header_bytes;       // sizeof(HgCkptHeader)
    uint32_t json_bytes;
    uint32_t map_w, map_h, row0, rows;
    int32_t erosion_type;
    uint32_t particle_count;
    uint32_t erosion_steps;
    uint32_t n_fields;
    int32_t field_ids[8];        // hg_field values, in file order
    hg_erosion_data erosion;     // 96 bytes
    hg_rain_data rain;           // 20 bytes
    uint32_t _pad0;
    hg_map_settings_data map;    // 96 bytes
    uint32_t _pad1;
    uint64_t payload_bytes;      // fields + particles (offset 304)
};
static_assert(sizeof(HgCkptHeader) == 312 && offsetof(HgCkptHeader, erosion) == 84 && offsetof(HgCkptHeader, map) == 204 && offsetof(HgCkptHeader, payload_bytes) == 304, "checkpoint header layout (hydro_gen_b200/checkpoint.py mirrors it)");

const char* field_name(int f) {
    switch (f) {
    case HG_FIELD_HEIGHTMAP: return "heightmap";
    case HG_FIELD_FLUX: return "flux";
    case HG_FIELD_VELOCITY: return "velocity";
    case HG_FIELD_SEDIMENT: return "sediment";
    default: return "?";
    }
}

std::string make_json(const HgCkptHeader& h) {
    char buf[2048];
    std::string fields;
    for (uint32_t k = 0; k < h.n_fields; k++) { fields += k ? ", \"" : "\""; fields += field_name(h.field_ids[k]); fields += "\""; }
    snprintf(buf, sizeof(buf),
             "{\"format\": \"hydrogen_b200 checkpoint\", \"version\": %u, \"map\": [%u, %u], \"row0\": %u, \"rows\": %u, "
             "\"erosion_type\": \"%s\", \"particle_count\": %u, \"erosion_steps\": %u, \"fields\": [%s], "
             "\"field_layout\": \"float32 little endian [row][x][rgba]\", "
             "\"erosion\": {\"Kc\": %.9g, \"Kalpha\": [%.9g, %.9g], \"Kconv\": %.9g, \"Ks\": [%.9g, %.9g], \"Kd\": [%.9g, %.9g], \"Ke\": %.9g, "
             "\"ENERGY_KEPT\": %.9g, \"Kspeed\": [%.9g, %.9g], \"G\": %.9g, \"d_t\": %.9g}, "
             "\"rain\": {\"amount\": %.9g, \"mountain_thresh\": %.9g, \"mountain_multip\": %.9g, \"period\": %d, \"drops\": %.9g}, "
             "\"map_settings\": {\"seed\": %.9g, \"max_height\": %.9g, \"max_dirt\": %.9g, \"scale\": %.9g, \"octaves\": %d}}",
             h.version, h.map_w, h.map_h, h.row0, h.rows, h.erosion_type == HG_GRID ? "grid" : "particles", h.particle_count, h.erosion_steps,
             fields.c_str(), h.erosion.Kc, h.erosion.Kalpha[0], h.erosion.Kalpha[1], h.erosion.Kconv, h.erosion.Ks[0], h.erosion.Ks[1],
             h.erosion.Kd[0], h.erosion.Kd[1], h.erosion.Ke, h.erosion.ENERGY_KEPT, h.erosion.Kspeed[0], h.erosion.Kspeed[1], h.erosion.G, h.erosion.d_t,
             h.rain.amount, h.rain.mountain_thresh, h.rain.mountain_multip, h.rain.period, h.rain.drops,
             h.map.seed, h.map.max_height, h.map.max_dirt, h.map.scale, h.map.octaves);
    return std::string(buf);
}

struct File {
    FILE* f = nullptr;
    ~File() { if (f) fclose(f); }
};

}  // namespace

extern "C" int hg_checkpoint_save(hg_ctx* c, const char* path) {
    HG_CHECK_CTX(c);
    if (!path) { hg_set_error("null checkpoint path"); return HG_ERR_INVALID; }
    HgCkptHeader h;
    memset(&h, 0, sizeof(h));
    memcpy(h.magic, "HGCKPT01", 8);
    h.version = 1; h.header_bytes = (uint32_t)sizeof(h);
    h.map_w = (uint32_t)c->g.W; h.map_h = (uint32_t)c->g.H; h.row0 = (uint32_t)c->g.row0; h.rows = (uint32_t)c->g.rows;
    h.erosion_type = c->erosion_type; h.particle_count = c->particle_count; h.erosion_steps = c->erosion_steps;
    // grid: V is dead between steps (hydro_flux.glsl:144-161 overwrites it); particles: V is the momentum map
    if (c->erosion_type == HG_GRID) { h.n_fields = 3; h.field_ids[0] = HG_FIELD_HEIGHTMAP; h.field_ids[1] = HG_FIELD_FLUX; h.field_ids[2] = HG_FIELD_SEDIMENT; }
    else { h.n_fields = 2; h.field_ids[0] = HG_FIELD_HEIGHTMAP; h.field_ids[1] = HG_FIELD_VELOCITY; }
    h.erosion = c->erosion; h.rain = c->rain; h.map = c->map;
    const size_t field_bytes = (size_t)c->g.rows * c->g.W * 4 * sizeof(float);
    const size_t part_bytes = c->erosion_type == HG_PARTICLES ? (size_t)c->particle_count * sizeof(hg_particle) : 0;
    h.payload_bytes = h.n_fields * field_bytes + part_bytes;
    const std::string json = make_json(h);
    h.json_bytes = (uint32_t)json.size();
    // written to a temporary file beside the target and renamed once complete and flushed, so a crash or a
    // full disk never leaves a half-written file under the checkpoint's name
    const std::string tmp = std::string(path) + ".tmp";
    File out;
    out.f = fopen(tmp.c_str(), "wb");
    if (!out.f) { hg_set_error("cannot open %s for writing", tmp.c_str()); return HG_ERR_INVALID; }
    int rc = HG_OK;
    if (fwrite(&h, sizeof(h), 1, out.f) != 1 || fwrite(json.data(), 1, json.size(), out.f) != json.size()) { hg_set_error("short write to %s", tmp.c_str()); rc = HG_ERR_STATE; }
    float* host = rc == HG_OK ? static_cast<float*>(hg_host_alloc(field_bytes > part_bytes ? field_bytes : part_bytes)) : nullptr;
    if (rc == HG_OK && !host) rc = HG_ERR_CUDA;
    for (uint32_t k = 0; k < h.n_fields && rc == HG_OK; k++) {
        rc = hg_download(c, h.field_ids[k], host);
        if (rc == HG_OK && fwrite(host, 1, field_bytes, out.f) != field_bytes) { hg_set_error("short write to %s", tmp.c_str()); rc = HG_ERR_STATE; }
    }
    if (rc == HG_OK && part_bytes) {
        rc = hg_download_particles(c, reinterpret_cast<hg_particle*>(host), c->particle_count);
        if (rc == HG_OK && fwrite(host, 1, part_bytes, out.f) != part_bytes) { hg_set_error("short write to %s", tmp.c_str()); rc = HG_ERR_STATE; }
    }
    if (host) hg_host_free(host);
    if (rc == HG_OK && fflush(out.f) != 0) { hg_set_error("flush of %s failed", tmp.c_str()); rc = HG_ERR_STATE; }
    FILE* f = out.f;
    out.f = nullptr;
    if (fclose(f) != 0 && rc == HG_OK) { hg_set_error("close of %s failed", tmp.c_str()); rc = HG_ERR_STATE; }
    if (rc == HG_OK && rename(tmp.c_str(), path) != 0) { hg_set_error("cannot rename %s to %s", tmp.c_str(), path); rc = HG_ERR_STATE; }
    if (rc != HG_OK) remove(tmp.c_str());
    return rc;
}

extern "C" int hg_checkpoint_load(hg_ctx* c, const char* path) {
    HG_CHECK_CTX(c);
    if (!path) { hg_set_error("null checkpoint path"); return HG_ERR_INVALID; }
    File in;
    in.f = fopen(path, "rb");
    if (!in.f) { hg_set_error("cannot open %s", path); return HG_ERR_INVALID; }
    HgCkptHeader h;
    if (fread(&h, sizeof(h), 1, in.f) != 1 || memcmp(h.magic, "HGCKPT01", 8) != 0 || h.version != 1 || h.header_bytes != sizeof(h) || h.n_fields > 8) {
        hg_set_error("%s is not a hydrogen_b200 checkpoint (version 1)", path); return HG_ERR_INVALID;
    }
    if ((int)h.map_w != c->g.W || (int)h.map_h != c->g.H || (int)h.row0 != c->g.row0 || (int)h.rows != c->g.rows ||
        h.erosion_type != c->erosion_type || h.particle_count != c->particle_count) {
        hg_set_error("checkpoint holds rows [%u,%u) of a %ux%u map (%s, %u droplets); this context holds rows [%d,%d) of %dx%d (%s, %u droplets)",
                     h.row0, h.row0 + h.rows, h.map_w, h.map_h, h.erosion_type == HG_GRID ? "grid" : "particles", h.particle_count,
                     c->g.row0, c->g.row0 + c->g.rows, c->g.W, c->g.H, c->erosion_type == HG_GRID ? "grid" : "particles", c->particle_count);
        return HG_ERR_INVALID;
    }
    const size_t field_bytes = (size_t)c->g.rows * c->g.W * 4 * sizeof(float);
    const size_t part_bytes = c->erosion_type == HG_PARTICLES ? (size_t)c->particle_count * sizeof(hg_particle) : 0;
    if (h.payload_bytes != h.n_fields * field_bytes + part_bytes) { hg_set_error("%s: payload size does not match its header", path); return HG_ERR_INVALID; }
    // the whole payload must be there BEFORE the context is touched: a truncated file leaves it as it was
    const long payload_at = (long)(sizeof(h) + h.json_bytes);
    if (fseek(in.f, 0, SEEK_END) != 0 || (uint64_t)ftell(in.f) < (uint64_t)payload_at + h.payload_bytes) { hg_set_error("%s is truncated", path); return HG_ERR_INVALID; }
    if (fseek(in.f, payload_at, SEEK_SET) != 0) { hg_set_error("cannot seek in %s", path); return HG_ERR_INVALID; }
    float* host = static_cast<float*>(hg_host_alloc(field_bytes > part_bytes ? field_bytes : part_bytes));
    if (!host) return HG_ERR_CUDA;
    int rc = hg_set_erosion(c, &h.erosion);
    if (rc == HG_OK) rc = hg_set_rain(c, &h.rain);
    if (rc == HG_OK) rc = hg_set_map(c, &h.map);
    for (uint32_t k = 0; k < h.n_fields && rc == HG_OK; k++) {
        if (fread(host, 1, field_bytes, in.f) != field_bytes) { hg_set_error("%s is truncated", path); rc = HG_ERR_INVALID; break; }
        rc = hg_upload(c, h.field_ids[k], host);
    }
    if (rc == HG_OK && part_bytes) {
        if (fread(host, 1, part_bytes, in.f) != part_bytes) { hg_set_error("%s is truncated", path); rc = HG_ERR_INVALID; }
        else rc = hg_upload_particles(c, reinterpret_cast<const hg_particle*>(host), c->particle_count);
    }
    if (rc == HG_OK) c->erosion_steps = h.erosion_steps;
    hg_host_free(host);
    // A connected slab also needs its neighbours' edge rows in its ghost rows (the file holds owned rows only):
    // collective push + wait, every rank of the slab table loads its own file.
    if (rc == HG_OK) rc = hg_slab_refresh_halo(c);
    return rc;
}

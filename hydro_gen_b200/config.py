"""config.ini of the reference (src/main.cpp:203-234): keys [window] width,height,
[map] size, [erosion] type, particle_count; written with defaults when missing."""
import configparser
import os

DEFAULT_TEXT = ("[window]\nwidth = 1280\nheight = 720\n\n[map]\nsize=1024\n\n[erosion]\n"
                "; type = grid or type = particle\ntype = grid\n"
                "; particle_count works only when the erosion type is \"particle\"\nparticle_count = 262144")


def load(path="config.ini", create=True):
    """Returns dict(window_w, window_h, map_size, erosion_type ('grid'|'particle'), particle_count)."""
    if not os.path.exists(path):
        if not create:
            raise FileNotFoundError(path)
        with open(path, "w") as f:
            f.write(DEFAULT_TEXT)
    cp = configparser.ConfigParser(inline_comment_prefixes=(";",), comment_prefixes=(";", "#"))
    cp.read(path)
    etype = cp.get("erosion", "type", fallback="grid").strip()
    particle = etype == "particle"     # anything else means grid (main.cpp:229-234)
    return {
        "window_w": cp.getint("window", "width", fallback=1280),
        "window_h": cp.getint("window", "height", fallback=720),
        "map_size": cp.getint("map", "size", fallback=1024),
        "erosion_type": "particle" if particle else "grid",
        "particle_count": cp.getint("erosion", "particle_count", fallback=262144) if particle else 0,
    }

"""Second, independent restatement of the reference's grid shaders in scalar numpy float32.

TEST INFRASTRUCTURE (known-answer generator).  Written from the GLSL text alone, cell by
cell, with Python loops; it shares no code with oracle/hg_oracle.c or with the CUDA
kernels.  tests/golden/make_golden.py runs it on small seeded cases and commits the
inputs and outputs as tests/golden/grid_cases.npz; the oracle (CPU tests) and the CUDA
path (GPU tests) must both reproduce those files bit for bit.  Two restatements written
separately from the same shader text agreeing to the last bit was the first pin of the
oracle; since then the reference's own shaders have been compiled for the CPU
(oracle/refshader/) and reproduce the files written by this module bit for bit
(tests/test_refshaders.py).

Every numpy float32 scalar operation is one correctly rounded IEEE operation, so the
operand order below IS the arithmetic.  Images are float32 arrays [y][x][4].

GLSL built-ins the specification leaves open are taken as include/hg_defined_math.h
defines them for this boundary (min/max by their GLSL 4.60 §8.3 formulas, smoothstep with
reversed edges, atan as the documented 3-range degree-9 polynomial); atan_defined below
restates that scheme and make_golden.py also checks it against numpy's arctan.
"""
import numpy as np

f32 = np.float32
ZERO, ONE = f32(0.0), f32(1.0)
L = f32(1.0)            # bindings.glsl:4  (WORLD_SCALE 1.0 -> L 1.0)
A_PIPE = f32(1.0)       # hydro_flux.glsl:31
OOB_HEIGHT = f32(999999999999.0)   # hydro_flux.glsl:37, thermal_erosion.glsl:24


def gmax(x, y):         # GLSL max(x, y) = (x < y) ? y : x
    return y if x < y else x


def gmin(x, y):         # GLSL min(x, y) = (y < x) ? y : x
    return y if y < x else x


def clamp(x, lo, hi):
    return gmin(gmax(x, lo), hi)


def mix(x, y, a):       # x*(1-a) + y*a
    return x * (ONE - a) + y * a


def fract(x):
    return x - np.floor(x)


def smoothstep(e0, e1, x):
    t = clamp((x - e0) / (e1 - e0), ZERO, ONE)
    return t * t * (f32(3.0) - f32(2.0) * t)


def atan_defined(xx):
    """atan as the boundary defines it (include/hg_defined_math.h): |x| reduced over three
    ranges (> tan(3pi/8): pi/2 - atan(1/x); > tan(pi/8): pi/4 + atan((x-1)/(x+1))), then the
    odd degree-9 polynomial x + x*z*(((c4 z + c3) z + c2) z + c1), z = x*x."""
    x = abs(xx)
    if x > f32(2.414213562373095):
        y = f32(1.5707963267948966)
        x = -(ONE / x)
    elif x > f32(0.4142135623730950):
        y = f32(0.7853981633974483)
        x = (x - ONE) / (x + ONE)
    else:
        y = ZERO
    z = x * x
    p = (((f32(8.05374449538e-2) * z - f32(1.38776856032e-1)) * z + f32(1.99777106478e-1)) * z
         - f32(3.33329491539e-1)) * z * x + x
    y = y + p
    return -y if xx < ZERO else y


class Params:
    """The members of Erosion_data (bindings.glsl:39-60) the grid shaders read."""

    def __init__(self, **kw):
        d = dict(Kc=0.2, Kalpha=(1.3, 0.6), Kconv=0.001, Ks=(0.03, 0.09), Kd=(0.01, 0.03), Ke=0.03,
                 ENERGY_KEPT=1.0, Kspeed=(0.5, 2.0), G=1.0, d_t=0.001, particle_count=0)   # state.cpp:81-92
        d.update(kw)
        self.Kc, self.Kconv, self.Ke = f32(d["Kc"]), f32(d["Kconv"]), f32(d["Ke"])
        self.ENERGY_KEPT, self.G, self.d_t = f32(d["ENERGY_KEPT"]), f32(d["G"]), f32(d["d_t"])
        self.Kalpha = [f32(v) for v in d["Kalpha"]]
        self.Ks = [f32(v) for v in d["Ks"]]
        self.Kd = [f32(v) for v in d["Kd"]]
        self.Kspeed = [f32(v) for v in d["Kspeed"]]
        self.particle_count = int(d["particle_count"])


def _oob(img, x, y):
    h, w = img.shape[:2]
    return x < 0 or x > w - 1 or y < 0 or y > h - 1


def fetch0(img, x, y):
    """texelFetch / imageLoad; out of bounds defined as 0 (SURVEY.md §8a hazard 3)."""
    if _oob(img, x, y):
        return np.zeros(4, f32)
    return img[y, x]


# ------------------------------------------------------------------ hydro_flux.glsl:77-166
def flux_pass(H, F, V, P):
    Ho, Fo, Vo = np.empty_like(H), np.empty_like(F), np.empty_like(V)
    hh, ww = H.shape[:2]

    def wheight(x, y):      # get_wheight, :33-39
        return OOB_HEIGHT if _oob(H, x, y) else H[y, x, 3]

    def flux(x, y):         # get_flux, :41-47
        return fetch0(F, x, y)

    with np.errstate(all="ignore"):
        for y in range(hh):
            for x in range(ww):
                out = flux(x, y).copy()
                vel = V[y, x].copy()
                terrain = H[y, x].copy()
                d1 = terrain[2]
                dh = [terrain[3] - wheight(x - 1, y), terrain[3] - wheight(x + 1, y),
                      terrain[3] - wheight(x, y + 1), terrain[3] - wheight(x, y - 1)]
                inf = [flux(x - 1, y)[1], flux(x + 1, y)[0], flux(x, y + 1)[3], flux(x, y - 1)[2]]
                for i in range(4):
                    out[i] = gmax(ZERO, P.ENERGY_KEPT * out[i] + P.d_t * A_PIPE * (P.G * dh[i]) / L)
                if x <= 0:
                    out[0] = ZERO
                elif x >= ww - 1:
                    out[1] = ZERO
                if y <= 0:
                    out[3] = ZERO
                elif y >= hh - 1:
                    out[2] = ZERO
                sum_in = inf[0] + inf[1] + inf[2] + inf[3]
                sum_out = out[0] + out[1] + out[2] + out[3]
                K = gmin(ONE, (terrain[2] * L * L) / (sum_out * P.d_t))
                for i in range(4):
                    out[i] = out[i] * K
                sum_out = sum_out * K
                d_volume = P.d_t * (sum_in - sum_out)
                d2 = gmax(ZERO, d1 + (d_volume / (L * L)))
                terrain[2] = d2
                terrain[3] = terrain[0] + d2 + terrain[1]
                vel[2] = d1 + d2
                if vel[2] > ZERO:
                    vel[0] = (flux(x - 1, y)[1] - flux(x, y)[0] + flux(x, y)[1] - flux(x + 1, y)[0]) / (L * vel[2])
                    vel[1] = (flux(x, y - 1)[2] - flux(x, y)[3] + flux(x, y)[2] - flux(x, y + 1)[3]) / (L * vel[2])
                else:
                    vel[0] = ZERO
                    vel[1] = ZERO
                Fo[y, x], Vo[y, x], Ho[y, x] = out, vel, terrain
    return Ho, Fo, Vo


# --------------------------------------------------------------- hydro_erosion.glsl:23-92
def erosion_pass(H, S, V, P):
    Ho, So = np.empty_like(H), np.empty_like(S)
    hh, ww = H.shape[:2]
    with np.errstate(all="ignore"):
        for y in range(hh):
            for x in range(ww):
                vel = V[y, x]
                terrain = H[y, x].copy()
                sediment = S[y, x].copy()
                dd = vel[2]
                length_v = np.sqrt(vel[0] * vel[0] + vel[1] * vel[1])
                if dd < f32(1e-3):
                    dd = gmax(f32(5e-4), dd)
                    ero_vel = mix(length_v, ZERO, smoothstep(f32(1e-3), f32(5e-4), dd))
                else:
                    ero_vel = length_v
                cap = ZERO
                # get_terr_normal :23-35
                r, l = fetch0(H, x + 1, y), fetch0(H, x - 1, y)
                b, t = fetch0(H, x, y - 1), fetch0(H, x, y + 1)
                dx = r[0] + r[1] - l[0] - l[1]
                dz = t[0] + t[1] - b[0] - b[1]
                a3 = (f32(2.0) * L, dx, ZERO)
                b3 = (ZERO, dz, f32(2.0) * L)
                cx = a3[1] * b3[2] - a3[2] * b3[1]
                cy = a3[2] * b3[0] - a3[0] * b3[2]
                cz = a3[0] * b3[1] - a3[1] * b3[0]
                inv = ONE / np.sqrt(cx * cx + cy * cy + cz * cz)
                ny = cy * inv
                sin_a = abs(np.sqrt(ONE - ny * ny))
                for i in (1, 0):
                    Kls = P.d_t * P.Ks[i]
                    Kld = P.d_t * P.Kd[i]
                    c = gmax(ZERO, P.Kc * gmax(f32(0.02), sin_a) * ero_vel - cap)
                    if c > sediment[i]:
                        old_terr = terrain[i]
                        delta = Kls * (c - sediment[i])
                        terrain[i] = terrain[i] - delta
                        sediment[i] = sediment[i] + delta
                        if terrain[i] < ZERO:
                            sediment[i] = sediment[i] + terrain[i]
                            terrain[i] = ZERO
                            cap = cap + old_terr
                        else:
                            break
                    else:
                        delta = Kld * (sediment[i] - c)
                        terrain[i] = terrain[i] + delta
                        sediment[i] = sediment[i] - delta
                conv = sediment[0] * P.Kconv * P.d_t
                sediment[1] = sediment[1] + conv
                sediment[0] = sediment[0] - conv
                terrain[3] = terrain[0] + terrain[1] + terrain[2]
                So[y, x], Ho[y, x] = sediment, terrain
    return Ho, So


# ------------------------------------- sediment_transport.glsl:66-93, img_interpolation.glsl
def img_bilinear(img, sx, sy):
    px, py = int(sx), int(sy)           # ivec2(): truncation
    fx, fy = fract(sx), fract(sy)
    v1 = [mix(a, b, fx) for a, b in zip(fetch0(img, px, py), fetch0(img, px + 1, py))]
    v2 = [mix(a, b, fx) for a, b in zip(fetch0(img, px, py + 1), fetch0(img, px + 1, py + 1))]
    return np.array([mix(a, b, fy) for a, b in zip(v1, v2)], f32)


def sediment_pass(H, S, V, P):
    Ho, So = np.empty_like(H), np.empty_like(S)
    hh, ww = H.shape[:2]
    with np.errstate(all="ignore"):
        for y in range(hh):
            for x in range(ww):
                vel = V[y, x]
                bx = f32(x) - vel[0] * P.d_t
                by = f32(y) - vel[1] * P.d_t
                bx = clamp(bx, ZERO, f32(ww - 1))
                by = clamp(by, ZERO, f32(hh - 1))
                if bx != bx:        # ivec2(NaN) is undefined in GLSL; defined as 0
                    bx = ZERO
                if by != by:
                    by = ZERO
                st = img_bilinear(S, bx, by)
                terrain = H[y, x].copy()
                terrain[2] = terrain[2] * (ONE - P.Ke * P.d_t)
                terrain[3] = terrain[0] + terrain[1] + terrain[2]
                So[y, x], Ho[y, x] = st, terrain
    return Ho, So


# ---------------------------------------------------------------- thermal_erosion.glsl:28-115
def thermal_outflow_pass(H, layer, P):
    TC, TD = np.zeros_like(H), np.zeros_like(H)
    hh, ww = H.shape[:2]
    offs = [[(-1, 0), (1, 0), (0, 1), (0, -1)], [(-1, 1), (1, 1), (-1, -1), (1, -1)]]

    def height(x, y):       # get_height :20-26
        return np.full(4, OOB_HEIGHT, f32) if _oob(H, x, y) else H[y, x]

    with np.errstate(all="ignore"):
        for y in range(hh):
            for x in range(ww):
                terrain = H[y, x]
                d_h = [[ZERO] * 4, [ZERO] * 4]
                for i in range(layer + 1):
                    for j in range(2):
                        for k in range(4):
                            ox, oy = offs[j][k]
                            d_h[j][k] = d_h[j][k] + (terrain[i] - height(x + ox, y + oy)[i])
                Hm = ZERO
                for j in range(2):
                    for k in range(4):
                        if d_h[j][k] > Hm:
                            Hm = d_h[j][k]
                Hm = gmin(terrain[layer], Hm)
                out = [[ZERO] * 4, [ZERO] * 4]
                bk = ZERO
                sharpness = ONE
                for j in range(2):
                    for k in range(4):
                        b = d_h[j][k]
                        if b <= ZERO:
                            continue
                        d = L
                        if j == 1:
                            d = d * np.sqrt(f32(2.0))
                        alph = atan_defined(b / (d / f32(1.0)))
                        if alph > P.Kalpha[layer]:
                            newsh = ONE + alph - P.Kalpha[layer]
                            if newsh > sharpness:
                                sharpness = newsh
                            bk = bk + b
                            out[j][k] = ONE
                sharpness = sharpness * (sharpness * sharpness)
                S = P.d_t * P.Kspeed[layer] * sharpness * L * Hm / f32(2.0)
                for j in range(2):
                    for k in range(4):
                        out[j][k] = S * d_h[j][k] / bk if out[j][k] == ONE else ZERO
                TC[y, x], TD[y, x] = out[0], out[1]
    return TC, TD


# ------------------------------------------------------------- thermal_transport.glsl:31-65
def thermal_transport_pass(H, TC, TD, layer):
    Ho = np.empty_like(H)
    hh, ww = H.shape[:2]
    for y in range(hh):
        for x in range(ww):
            in_flux = ZERO
            in_flux = in_flux + fetch0(TC, x - 1, y)[1]
            in_flux = in_flux + fetch0(TC, x + 1, y)[0]
            in_flux = in_flux + fetch0(TC, x, y + 1)[3]
            in_flux = in_flux + fetch0(TC, x, y - 1)[2]
            in_flux = in_flux + fetch0(TD, x - 1, y + 1)[3]
            in_flux = in_flux + fetch0(TD, x + 1, y + 1)[2]
            in_flux = in_flux + fetch0(TD, x - 1, y - 1)[1]
            in_flux = in_flux + fetch0(TD, x + 1, y - 1)[0]
            sum_flux = ZERO
            for img in (TC, TD):
                for k in range(4):
                    sum_flux = sum_flux - img[y, x, k]
            sum_flux = sum_flux + in_flux
            terrain = H[y, x].copy()
            terrain[layer] = terrain[layer] + sum_flux
            terrain[3] = terrain[0] + terrain[1] + terrain[2]
            Ho[y, x] = terrain
    return Ho


# -------------------------------------------------------------------- smoothing.glsl:22-103
def smooth_pass(H, P):
    """Grid mode (particle_count == 0): the momentum branch is not taken."""
    Ho = np.empty_like(H)
    hh, ww = H.shape[:2]
    for y in range(hh):
        for x in range(ww):
            terrain = H[y, x].copy()
            if x == 0 or y == 0 or x == ww - 1 or y == hh - 1:
                Ho[y, x] = terrain
                continue
            terr = [terrain[0], terrain[1]]
            l, r, t, b = H[y, x - 1], H[y, x + 1], H[y + 1, x], H[y - 1, x]

            def diff(n):
                d = [terr[0] - n[0], terr[1] - n[1]]
                d[1] = d[1] + d[0]
                return d
            d_l, d_r, d_t, d_b = diff(l), diff(r), diff(t), diff(b)
            g_hdiff = abs((d_l[1] + d_r[1] + d_t[1] + d_b[1]) / f32(4.0))
            r_hdiff = abs((d_l[0] + d_r[0] + d_t[0] + d_b[0]) / f32(4.0))
            x_crv = [d_l[0] * d_r[0], d_l[1] * d_r[1]]
            y_crv = [d_t[0] * d_b[0], d_t[1] * d_b[1]]
            if (((-d_l[0]) > r_hdiff or (-d_r[0]) > r_hdiff) and x_crv[0] > ZERO) or \
               (((-d_t[0]) > r_hdiff or (-d_b[0]) > r_hdiff) and y_crv[0] > ZERO):
                terr[0] = (terr[0] + l[0] + r[0] + t[0] + b[0]) / f32(5.0)
            if (((-d_l[1]) > g_hdiff or (-d_r[1]) > g_hdiff) and x_crv[1] > ZERO) or \
               (((-d_t[1]) > g_hdiff or (-d_b[1]) > g_hdiff) and y_crv[1] > ZERO):
                terr[1] = (terr[1] + l[1] + r[1] + t[1] + b[1]) / f32(5.0)
            multip = clamp(P.Kspeed[1] * P.d_t, ZERO, ONE)
            terrain[0] = multip * terr[0] + (ONE - multip) * terrain[0]
            terrain[1] = multip * terr[1] + (ONE - multip) * terrain[1]
            terrain[3] = terrain[0] + terrain[1] + terrain[2]
            Ho[y, x] = terrain
    return Ho


# ----------------------------------------------- Erosion::dispatch_grid, erosion.cpp:158-200
def grid_step(H, F, V, S, P, trace=None):
    """One dispatch_grid: returns (H, F, V, S, TC, TD) as the read textures hold them after
    the step.  `trace`, if a dict, receives the images after each of the 8 dispatches."""
    H, F, V = flux_pass(H, F, V, P)
    if trace is not None:
        trace["flux"] = (H.copy(), F.copy(), V.copy())
    H, S = erosion_pass(H, S, V, P)
    if trace is not None:
        trace["erosion"] = (H.copy(), S.copy())
    H, S = sediment_pass(H, S, V, P)
    if trace is not None:
        trace["sediment"] = (H.copy(), S.copy())
    TC = TD = None
    for layer in (0, 1):
        TC, TD = thermal_outflow_pass(H, layer, P)
        H = thermal_transport_pass(H, TC, TD, layer)
        if trace is not None:
            trace[f"thermal{layer}"] = (H.copy(), TC.copy(), TD.copy())
    H = smooth_pass(H, P)
    if trace is not None:
        trace["smooth"] = (H.copy(),)
    return H, F, V, S, TC, TD

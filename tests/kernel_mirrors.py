"""Line-by-line Python mirrors of the fold and frame kernels (packnet_sfm_b200/csrc/fold_kernels.cu, frame_kernels.cu):
thread -> element maps, shared-memory staging, store indices, loaders and the (a, l) -> (row, col) maps, restated with flat
arrays exactly as the kernels index them.  TEST INFRASTRUCTURE: the kernels only run on the GPU tier; the mirrors pin their
index arithmetic on the CPU tier (tests/test_fold_mirror_cpu.py, tests/test_frame_mirror_cpu.py) in plain Python, next to
the host emulation that executes the real sources (tests/emu/, tests/test_kernels_emulated_cpu.py).  The frame mirror states
the reduction as one flat index; the kernel walks the same index space as nested loops."""
import numpy as np

T = 128
def perm_channel(cpp, C): return (cpp & 3) * C + (cpp >> 2)

def fwd(w2, w3, cout, n, k, ky0, ky1, kx0, kx1, dy0, dy1, dx0, dx1, ohwi=0):
    KA, KB, DA, DB = ky1-ky0, kx1-kx0, dy1-dy0, dx1-dx0
    EA, EB = KA+DA-1, KB+DB-1; E = EA*EB; C = n//4
    w2f, w3f = w2.reshape(-1), w3.reshape(-1)
    out = np.full(cout*n*E, np.nan)
    nblk = (n+T-1)//T
    for co in range(cout):
        for bx in range(nblk):
            w3s = np.zeros(8*3*DA*DB)
            for i in range(8*3*DA*DB):
                dx = i % DB; dy = (i//DB) % DA; fdc = i//(DA*DB)
                w3s[i] = w3f[fdc*9 + (dy0+dy)*3 + (dx0+dx)]
            stage = np.zeros(T*E)
            c0 = bx*T
            for tid in range(T):
                cpp = c0+tid
                acc = np.zeros((EA, EB))
                if cpp < n:
                    for f in range(8):
                        for dc in range(3):
                            cp = cpp-dc+1
                            if cp < 0 or cp >= n: continue
                            src = ((((co*8+f)*n+cp)*k+ky0)*k+kx0)
                            v = np.array([[w2f[src+a*k+b] for b in range(KB)] for a in range(KA)])
                            wf = (f*3+dc)*DA*DB
                            for dy in range(DA):
                                for dx in range(DB):
                                    w = w3s[wf+dy*DB+dx]
                                    for a in range(KA):
                                        for b in range(KB):
                                            acc[a+dy][b+dx] += v[a][b]*w
                q = tid & 3; cl = tid >> 2
                for a in range(EA):
                    for b in range(EB):
                        stage[(q*32+cl)*E + a*EB+b] = acc[a][b]
            cbase = c0 >> 2
            nc = min(C-cbase, 32)
            for idx in range(4*32*E):
                if not ohwi:
                    qq = idx//(32*E); r = idx - qq*32*E
                    if r < nc*E:
                        out[(co*n + qq*C + cbase)*E + r] = stage[idx]
                else:
                    c = idx & 31; qq = (idx >> 5) & 3; e = idx >> 7
                    if c < nc:
                        out[(co*E + e)*n + qq*C + cbase + c] = stage[(qq*32 + c)*E + e]
    return out.reshape(cout, EA, EB, n) if ohwi else out.reshape(cout, n, EA, EB)

def bwd(w2, w3, dout, dS, dw2, dw3, accumulate, cout, n, k, ky0, ky1, kx0, kx1, dy0, dy1, dx0, dx1, ohwi=0):
    KA, KB, DA, DB = ky1-ky0, kx1-kx0, dy1-dy0, dx1-dx0
    EA, EB = KA+DA-1, KB+DB-1; E = EA*EB; C = n//4
    w2f, w3f, df = w2.reshape(-1), w3.reshape(-1), dout.reshape(-1)
    dw2f, dw3f = dw2.reshape(-1), dw3.reshape(-1)
    dSf = dS.reshape(-1) if dS is not None else None
    nblk = (n+T-1)//T
    for co in range(cout):
        for f in range(8):
            for bx in range(nblk):
                w3s = np.zeros(3*DA*DB); red = np.zeros(3*DA*DB)
                for tid in range(3*DA*DB):
                    dx = tid % DB; dy = (tid//DB) % DA; dc = tid//(DA*DB)
                    w3s[tid] = w3f[(f*3+dc)*9 + (dy0+dy)*3 + (dx0+dx)]
                for tid in range(T):
                    cp = bx*T+tid
                    if cp >= n: continue
                    g2 = np.zeros((KA,KB)); g3 = np.zeros((3,DA,DB))
                    off = ((((co*8+f)*n+cp)*k+ky0)*k+kx0)
                    v = np.array([[w2f[off+a*k+b] for b in range(KB)] for a in range(KA)])
                    for dc in range(3):
                        cpp = cp+dc-1
                        if cpp < 0 or cpp >= n: continue
                        pc = perm_channel(cpp, C)
                        d = co*E*n + pc if ohwi else (co*n + pc)*E
                        es = n if ohwi else 1
                        for ea in range(EA):
                            for eb in range(EB):
                                dv = df[d+(ea*EB+eb)*es]
                                for dy in range(DA):
                                    for dx in range(DB):
                                        a = ea-dy; b = eb-dx
                                        if 0 <= a < KA and 0 <= b < KB:
                                            g2[a][b] += dv*w3s[(dc*DA+dy)*DB+dx]
                                            g3[dc][dy][dx] += dv*v[a][b]
                    for a in range(KA):
                        for b in range(KB):
                            r = g2[a][b]
                            if dSf is not None: r += dSf[((co*8+f)*k+ky0)*k+kx0 + a*k+b]
                            if accumulate: r += dw2f[off+a*k+b]
                            dw2f[off+a*k+b] = r
                    red += g3.reshape(-1)
                for tid in range(3*DA*DB):
                    dx = tid % DB; dy = (tid//DB) % DA; dc = tid//(DA*DB)
                    dw3f[(f*3+dc)*9+(dy0+dy)*3+(dx0+dx)] += red[tid]



# ---------------------------------------------------------------------------------------------------------------
# frame_kernels.cu
# ---------------------------------------------------------------------------------------------------------------
def border_class(p, ln, m):
    return p if p < m else (p - (ln - m) + m + 1 if p >= ln - m else m)


def zmap(T, a, l):
    a1, a2 = a // T["A2"], a % T["A2"]
    return T["r0"] + a1*T["ra1"] + a2*T["ra2"] + l*T["rl"], T["c0"] + a1*T["ca1"] + a2*T["ca2"] + l*T["cl"]


def widx(T, co, a, e, nn):
    return co*T["w_sco"] + a*T["w_sa"] + e*T["w_se"] + nn


def load_line(T, n, b, l, e, nn):
    """T["line"]: flat array starting at the term's first pixel; batch stride T["line_bs"]."""
    pos = l + e - T["pad"]
    if pos < 0 or pos >= T["L"]:
        return 0.0
    return T["line"][b*T["line_bs"] + pos*n + nn]


def frame_forward(terms, B, h, w, Co, n, m, dB, z):
    """terms: dicts with the fields of pn_frame_term (line / w / dline / dw = flat numpy arrays); z [B,h,w,Co] accumulated"""
    g = 2*m + 1
    for T in terms:
        M, N, K = B*T["L"], T["A"]*Co, T["KE"]*n
        for p in range(M):
            b, l = p // T["L"], p % T["L"]
            for c in range(N):
                a, co = c // Co, c % Co
                acc = 0.0
                for kk in range(K):
                    e, nn = kk // n, kk % n
                    acc += load_line(T, n, b, l, e, nn) * T["w"][widx(T, co, a, e, nn)]
                row, col = zmap(T, a, l)
                v = T["alpha"]*acc
                if T["bias_mode"] == 1 or (T["bias_mode"] == 2 and m <= row < h - m):
                    v += dB[(border_class(row, h, m)*g + border_class(col, w, m))*Co + co]
                z[((b*h + row)*w + col)*Co + co] += v


def frame_backward(terms, B, h, w, Co, n, m, gz, gdB):
    g = 2*m + 1
    for T in terms:
        M, N, K = B*T["L"], n, T["A"]*T["KE"]*Co
        for p in range(M):
            b, j = p // T["L"], p % T["L"]
            for nn in range(N):
                acc = 0.0
                for kk in range(K):
                    co, ae = kk % Co, kk // Co
                    e, a = ae % T["KE"], ae // T["KE"]
                    l = j - e + T["pad"]
                    if l < 0 or l >= T["L"]:
                        continue
                    row, col = zmap(T, a, l)
                    acc += gz[((b*h + row)*w + col)*Co + co] * T["w"][widx(T, co, a, e, nn)]
                T["dline"][b*T["dline_bs"] + j*n + nn] += T["alpha"]*acc
        Mw, Nw, Kw = T["A"]*Co, T["KE"]*n, B*T["L"]
        for r in range(Mw):
            a, co = r // Co, r % Co
            for c in range(Nw):
                e, nn = c // n, c % n
                acc = 0.0
                for p in range(Kw):
                    b, l = p // T["L"], p % T["L"]
                    row, col = zmap(T, a, l)
                    acc += gz[((b*h + row)*w + col)*Co + co] * load_line(T, n, b, l, e, nn)
                T["dw"][widx(T, co, a, e, nn)] = T["alpha"]*acc
    F_ = h*w - (h - 2*m)*(w - 2*m)
    for b in range(B):
        for f in range(F_):
            if f < m*w:
                row, col = f // w, f % w
            elif f < 2*m*w:
                q = f - m*w
                row, col = h - m + q // w, q % w
            else:
                q = f - 2*m*w
                row = m + q // (2*m)
                cc = q % (2*m)
                col = cc if cc < m else w - 2*m + cc
            for co in range(Co):
                gdB[(border_class(row, h, m)*g + border_class(col, w, m))*Co + co] += gz[((b*h + row)*w + col)*Co + co]

"""Line-by-line Python model of direct2d_tiled_kernel (csrc/spatial_smooth.cu): same box staging, same loops, same
index arithmetic, one Python iteration per CUDA thread.  TEST TOOLING: tests/test_host_logic.py checks it against the
oracle convolution, which is how the kernel's indexing was verified before it could run on hardware."""
import numpy as np
TX, RY = 32, 8


def tiled(img, K, nw=8):
    """nw = warps per CTA (8 or 4): the tile is 32 columns x 8 nw rows."""
    TY = nw * RY
    ny, nx = img.shape; nty, ntx = K.shape; hy, hx = nty >> 1, ntx >> 1
    Kn = K / K.sum()
    bw, bh = TX + ntx - 1, TY + nty - 1 + RY
    out = np.full((ny, nx), -777.0)
    for ty_i in range(-(-ny // TY)):
        for tx_i in range(-(-nx // TX)):
            x0, y0 = tx_i * TX, ty_i * TY
            sval = np.zeros(bw * bh); sok = np.zeros(bw * bh, np.float32)
            for i in range(bw * bh):
                q, pcol = divmod(i, bw)
                yy, xx = y0 - hy + q, x0 - hx + pcol
                v = np.float32(0.0)
                if q < TY + nty - 1 and 0 <= xx < nx and 0 <= yy < ny:
                    v = img[yy, xx]
                good = v == v
                sval[i] = float(v) if good else 0.0; sok[i] = 1.0 if good else 0.0
            for warp in range(nw):
                r0 = warp * RY
                for lane in range(32):
                    top = np.zeros(RY); botd = np.zeros(RY)
                    for sx in range(ntx):
                        kx = ntx - 1 - sx
                        base = r0 * bw + lane + sx
                        a = np.zeros(2 * RY); o = np.zeros(2 * RY, np.float32); botf = np.zeros(RY, np.float32)
                        for j in range(RY):
                            a[j] = sval[base + j * bw]; o[j] = sok[base + j * bw]
                        for sb in range(0, nty, RY):
                            for j in range(RY):
                                a[RY + j] = sval[base + (sb + RY + j) * bw]; o[RY + j] = sok[base + (sb + RY + j) * bw]
                            for si in range(RY):
                                if sb + si < nty:
                                    k = Kn[nty - 1 - sb - si, kx]; kf = np.float32(k)
                                    for j in range(RY):
                                        top[j] += k * a[j + si]; botf[j] = np.float32(botf[j] + kf * o[j + si])
                            a[:RY] = a[RY:]; o[:RY] = o[RY:]
                        botd += botf.astype(np.float64)
                    x = x0 + lane
                    for j in range(RY):
                        y = y0 + r0 + j
                        if x < nx and y < ny:
                            ci = (r0 + j + hy) * bw + lane + hx
                            centre = sval[ci] if sok[ci] != 0 else np.nan
                            out[y, x] = centre if botd[j] == 0 else top[j] / botd[j]
    return out

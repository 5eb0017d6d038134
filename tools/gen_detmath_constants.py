"""Regenerates the constants in hanamaru_renderer_b200/csrc/hnm_detmath.h (mpmath, 400 bits)."""
import mpmath as mp
mp.mp.prec = 400

def R(z):
    z = mp.mpf(z)
    if z == 0:
        return mp.mpf(1) / 6
    x = mp.sqrt(z)
    return (mp.asin(x) / x - 1) / z

if __name__ == "__main__":
    for n in (10, 12, 13, 14, 16):
        poly, err = mp.chebyfit(R, [0, mp.mpf(1) / 4], n, error=True)
        print(n, mp.nstr(err, 5))
    n = 14
    poly = mp.chebyfit(R, [0, mp.mpf(1) / 4], n)
    # round to double and measure the error of the rounded polynomial
    coef = [float(c) for c in poly]
    worst = 0
    for i in range(2001):
        z = mp.mpf(i) / 2000 / 4
        acc = mp.mpf(0)
        for c in coef:
            acc = acc * z + mp.mpf(c)
        worst = max(worst, abs(acc - R(z)))
    print("rounded-poly max abs err", mp.nstr(worst, 5))
    print("// highest degree first")
    for c in coef:
        print("   ", c.hex())

"""Generate the 32-entry table of glibc's expf (sysdeps/ieee754/flt-32/e_exp2f_data.c, the
ARM optimized-routines algorithm, glibc >= 2.27):  T[i] = asuint64(2^(i/32)) - (i << 47),
with 2^(i/32) correctly rounded to double.  Computed from first principles with exact rational
arithmetic (bisection on integers), so nothing is copied from glibc."""
from fractions import Fraction
import struct

def cr_double_of_root(i, n=32):
    # correctly rounded double of 2^(i/n), 0 <= i < n: value in [1,2) => mantissa m/2^52, m in [2^52, 2^53)
    # find largest m with (m/2^52)^n <= 2^i, then round to nearest by testing the midpoint
    lo, hi = 1 << 52, 1 << 53
    tgt = Fraction(2) ** i
    while hi - lo > 1:
        mid = (lo + hi) // 2
        if Fraction(mid, 1 << 52) ** n <= tgt:
            lo = mid
        else:
            hi = mid
    midpoint = Fraction(2 * lo + 1, 1 << 53)
    m = lo + 1 if midpoint ** n <= tgt else lo  # ties impossible (irrational) except i = 0
    return m / float(1 << 52)

if __name__ == "__main__":
    for i in range(32):
        d = cr_double_of_root(i)
        bits = struct.unpack("<Q", struct.pack("<d", d))[0]
        t = (bits - (i << 47)) & 0xFFFFFFFFFFFFFFFF
        print("0x%016xULL," % t, end="\n" if i % 4 == 3 else " ")

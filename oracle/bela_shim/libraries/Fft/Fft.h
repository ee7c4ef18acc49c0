// Stand-in for Bela's libraries/Fft/Fft.h (NE10-backed on the board; not vendored by the
// reference, upmix.cpp:15).  TEST INFRASTRUCTURE ONLY.  Interface as used by upmix.cpp:194-285:
// setup(n) -> 0 on success, td(n) time-domain sample, fdr(k)/fdi(k) real/imag of bin k,
// fft() real -> half spectrum (unscaled), ifft() half spectrum -> real.  The inverse is scaled by
// 1/n: that is the NE10 r2c/c2r convention ASSUMED here (the reference does not show it); it is
// the scaling under which upmix.cpp's overlap-add of BH*BH windows has gain ~1.03.
// Arithmetic is double-precision radix-2 so the shim adds no error of its own.
#pragma once
#include <cmath>
#include <complex>
#include <vector>

class Fft {
public:
    int setup(unsigned int n) {
        if (n == 0 || (n & (n - 1))) return -1;
        n_ = n;
        td_.assign(n, 0.f);
        re_.assign(n / 2 + 1, 0.f);
        im_.assign(n / 2 + 1, 0.f);
        work_.assign(n, std::complex<double>(0, 0));
        return 0;
    }
    float& td(unsigned int i) { return td_[i]; }
    float& fdr(unsigned int k) { return re_[k]; }
    float& fdi(unsigned int k) { return im_[k]; }
    void fft() {
        for (unsigned int i = 0; i < n_; i++) work_[i] = std::complex<double>(td_[i], 0.0);
        transform(-1.0);
        for (unsigned int k = 0; k <= n_ / 2; k++) {
            re_[k] = (float)work_[k].real();
            im_[k] = (float)work_[k].imag();
        }
    }
    void ifft() {
        work_[0] = std::complex<double>(re_[0], 0.0);
        work_[n_ / 2] = std::complex<double>(re_[n_ / 2], 0.0);
        for (unsigned int k = 1; k < n_ / 2; k++) {
            work_[k] = std::complex<double>(re_[k], im_[k]);
            work_[n_ - k] = std::complex<double>(re_[k], -im_[k]);
        }
        transform(+1.0);
        for (unsigned int i = 0; i < n_; i++) td_[i] = (float)(work_[i].real() / (double)n_);
    }
private:
    void transform(double sign) {
        const unsigned int n = n_;
        for (unsigned int i = 1, j = 0; i < n; i++) {
            unsigned int bit = n >> 1;
            for (; j & bit; bit >>= 1) j ^= bit;
            j ^= bit;
            if (i < j) std::swap(work_[i], work_[j]);
        }
        for (unsigned int len = 2; len <= n; len <<= 1) {
            const double ang = sign * 2.0 * M_PI / (double)len;
            for (unsigned int i = 0; i < n; i += len) {
                for (unsigned int k = 0; k < len / 2; k++) {
                    const std::complex<double> w(std::cos(ang * k), std::sin(ang * k));
                    const std::complex<double> u = work_[i + k], v = work_[i + k + len / 2] * w;
                    work_[i + k] = u + v;
                    work_[i + k + len / 2] = u - v;
                }
            }
        }
    }
    unsigned int n_ = 0;
    std::vector<float> td_, re_, im_;
    std::vector<std::complex<double>> work_;
};

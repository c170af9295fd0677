// unipoly.hpp -- host mirror of [ARK] ark_poly::univariate::SparsePolynomial<F>, the message type
// of SumCheckPolynomial::to_univariate (/root/reference/sum-check-protocol/src/lib.rs:148).
// The zero-coefficient conventions decide transcript bytes (hard part 2) and are
// implemented literally: from_coefficients_vec pops trailing zero terms in the given order, then
// sorts; Add merges sorted term lists and drops an equal-degree pair only when it sums to zero;
// adding a zero polynomial returns the other operand unchanged; Dense->Sparse keeps exactly the
// non-zero coefficients.
#pragma once
#include <algorithm>
#include <array>
#include <cstdint>
#include <map>
#include <memory>
#include <mutex>
#include <utility>
#include <vector>

#include "hostfield.hpp"

namespace scb {

struct SparsePoly {
    std::vector<std::pair<uint64_t, Fe>> coeffs;  // (degree, coefficient), ascending degree

    static SparsePoly zero() { return SparsePoly(); }

    static SparsePoly from_coefficients_vec(const HostField& F, std::vector<std::pair<uint64_t, Fe>> c) {
        while (!c.empty() && F.is_zero(c.back().second)) c.pop_back();
        std::stable_sort(c.begin(), c.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
        SparsePoly p;
        p.coeffs = std::move(c);
        return p;
    }
    // [ARK] impl From<DensePolynomial> for SparsePolynomial
    static SparsePoly from_dense(const HostField& F, const std::vector<Fe>& dense) {
        std::vector<std::pair<uint64_t, Fe>> c;
        for (size_t i = 0; i < dense.size(); ++i)
            if (!F.is_zero(dense[i])) c.emplace_back((uint64_t)i, dense[i]);
        return from_coefficients_vec(F, std::move(c));
    }
    bool is_zero(const HostField& F) const {
        for (const auto& t : coeffs)
            if (!F.is_zero(t.second)) return false;
        return true;
    }
    SparsePoly add(const HostField& F, const SparsePoly& o) const {
        if (is_zero(F)) return o;
        if (o.is_zero(F)) return *this;
        SparsePoly r;
        size_t i = 0, j = 0;
        while (true) {
            if (i == coeffs.size() && j == o.coeffs.size()) break;
            if (i == coeffs.size()) {
                r.coeffs.insert(r.coeffs.end(), o.coeffs.begin() + j, o.coeffs.end());
                break;
            }
            if (j == o.coeffs.size()) {
                r.coeffs.insert(r.coeffs.end(), coeffs.begin() + i, coeffs.end());
                break;
            }
            const auto& a = coeffs[i];
            const auto& b = o.coeffs[j];
            if (a.first < b.first) {
                r.coeffs.push_back(a);
                ++i;
            } else if (a.first == b.first) {
                Fe s = F.add(a.second, b.second);
                if (!F.is_zero(s)) r.coeffs.emplace_back(a.first, s);
                ++i;
                ++j;
            } else {
                r.coeffs.push_back(b);
                ++j;
            }
        }
        return r;
    }
    // [ARK] Polynomial::evaluate: sum of c * x^deg
    Fe evaluate(const HostField& F, const Fe& x) const {
        Fe acc = F.zero();
        if (is_zero(F)) return acc;
        for (const auto& t : coeffs) acc = F.add(acc, F.mul(t.second, F.pow_u64(x, t.first)));
        return acc;
    }
    // [ARK] CanonicalSerialize: u64 LE length, then per term u64 LE degree + canonical coefficient
    void serialize(const HostField& F, std::vector<uint8_t>& out) const {
        put_u64(out, coeffs.size());
        for (const auto& t : coeffs) {
            put_u64(out, t.first);
            F.serialize(t.second, out);
        }
    }
    // returns bytes consumed, 0 on malformed input
    static size_t deserialize(const HostField& F, const uint8_t* data, size_t len, SparsePoly& out) {
        if (len < 8) return 0;
        uint64_t n = get_u64(data);
        size_t off = 8;
        const size_t eb = F.ser_bytes();
        if (n > (len - off) / (8 + eb)) return 0;
        out.coeffs.clear();
        for (uint64_t i = 0; i < n; ++i) {
            uint64_t d = get_u64(data + off);
            off += 8;
            Fe c;
            if (!F.deserialize(data + off, c)) return 0;
            off += eb;
            out.coeffs.emplace_back(d, c);
        }
        return off;
    }
    static void put_u64(std::vector<uint8_t>& out, uint64_t v) {
        for (int i = 0; i < 8; ++i) out.push_back((uint8_t)(v >> (8 * i)));
    }
    static uint64_t get_u64(const uint8_t* d) {
        uint64_t v = 0;
        for (int i = 0; i < 8; ++i) v |= (uint64_t)d[i] << (8 * i);
        return v;
    }
};

// /root/reference/matrix-multiplication/src/lib.rs:17-60, literally: three scaled basis polynomials
// built with from_coefficients_vec (so the 2nd and 3rd carry an explicit (0, 0) term for x0 = 0)
// and summed with SparsePolynomial::add.  9 field divisions, as in the reference.
inline SparsePoly interpolate_quadratic_poly(const HostField& F, const Fe (&x)[3], const Fe (&y)[3]) {
    auto basis = [&](int a, int b, int c) {  // basis polynomial of point a w.r.t. points b, c
        Fe den = F.mul(F.sub(x[a], x[b]), F.sub(x[a], x[c]));
        Fe den_inv = F.inverse(den);
        std::vector<std::pair<uint64_t, Fe>> co;
        co.emplace_back(0, F.mul(x[b], x[c]));
        co.emplace_back(1, F.sub(F.neg(x[b]), x[c]));
        co.emplace_back(2, F.one());
        for (auto& t : co) t.second = F.mul(F.mul(t.second, y[a]), den_inv);
        return SparsePoly::from_coefficients_vec(F, std::move(co));
    };
    SparsePoly p1 = basis(0, 1, 2), p2 = basis(1, 0, 2), p3 = basis(2, 0, 1);
    return p1.add(F, p2).add(F, p3);
}

// Per-field constants of the two interpolations, computed once (the field inversions are the expensive part:
// a Fermat inversion is ~380 multiplications in a 255-bit field).
struct InterpConsts {
    Fe x[3];           // 0, 1, 2
    Fe den_inv[3];     // 1/((x_a-x_b)(x_a-x_c)) for the three Lagrange basis polynomials on {0,1,2}
    std::vector<std::vector<Fe>> basis[10];  // basis[n][i][k] = coefficient k of the i-th Lagrange basis on {0..n-1}
};
inline const InterpConsts& interp_consts(const HostField& F) {
    static std::mutex mu;
    static std::map<std::array<uint64_t, kHostMaxLimbs>, std::unique_ptr<InterpConsts>> cache;
    std::array<uint64_t, kHostMaxLimbs> key{{F.p[0], F.p[1], F.p[2], F.p[3]}};
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(key);
    if (it != cache.end()) return *it->second;
    auto c = std::make_unique<InterpConsts>();
    c->x[0] = F.zero();
    c->x[1] = F.one();
    c->x[2] = F.add(F.one(), F.one());
    const int perm[3][3] = {{0, 1, 2}, {1, 0, 2}, {2, 0, 1}};
    for (int a = 0; a < 3; ++a) {
        Fe den = F.mul(F.sub(c->x[perm[a][0]], c->x[perm[a][1]]), F.sub(c->x[perm[a][0]], c->x[perm[a][2]]));
        c->den_inv[a] = F.inverse(den);
    }
    for (size_t n = 1; n < 10; ++n) {
        if (F.n == 1 && (uint64_t)n > F.p[0]) break;  // points 0..n-1 must be distinct mod p
        c->basis[n].assign(n, std::vector<Fe>());
        for (size_t i = 0; i < n; ++i) {
            std::vector<Fe> num(1, F.one());
            Fe den = F.one();
            for (size_t j = 0; j < n; ++j) {
                if (j == i) continue;
                Fe fj = F.from_u64(j);
                std::vector<Fe> nn(num.size() + 1, F.zero());
                for (size_t k = 0; k < num.size(); ++k) {
                    nn[k + 1] = F.add(nn[k + 1], num[k]);
                    nn[k] = F.sub(nn[k], F.mul(fj, num[k]));
                }
                num.swap(nn);
                den = F.mul(den, F.sub(F.from_u64(i), fj));
            }
            Fe dinv = F.inverse(den);
            for (auto& v : num) v = F.mul(v, dinv);
            c->basis[n][i] = num;
        }
    }
    auto* raw = c.get();
    cache[key] = std::move(c);
    return *raw;
}

// interpolate_quadratic_poly on the fixed points X = 0, 1, 2 with the three denominators' inverses cached:
// same terms, same explicit zeros, same merge order as the literal function above.
inline SparsePoly interpolate_quadratic_012(const HostField& F, const InterpConsts& c, const Fe (&y)[3]) {
    auto basis = [&](int a, int b, int cc) {
        std::vector<std::pair<uint64_t, Fe>> co;
        co.emplace_back(0, F.mul(c.x[b], c.x[cc]));
        co.emplace_back(1, F.sub(F.neg(c.x[b]), c.x[cc]));
        co.emplace_back(2, F.one());
        for (auto& t : co) t.second = F.mul(F.mul(t.second, y[a]), c.den_inv[a]);
        return SparsePoly::from_coefficients_vec(F, std::move(co));
    };
    SparsePoly p1 = basis(0, 1, 2), p2 = basis(1, 0, 2), p3 = basis(2, 0, 1);
    return p1.add(F, p2).add(F, p3);
}

// Same coefficients as lagrange_to_coeffs below, from the cached basis (sum_i y_i * basis_i).
inline std::vector<Fe> lagrange_to_coeffs_cached(const HostField& F, const InterpConsts& c, const std::vector<Fe>& ys) {
    const size_t n = ys.size();
    std::vector<Fe> coeffs(n, F.zero());
    const auto& B = c.basis[n];
    for (size_t i = 0; i < n; ++i)
        for (size_t k = 0; k < n; ++k) coeffs[k] = F.add(coeffs[k], F.mul(ys[i], B[i][k]));
    while (!coeffs.empty() && F.is_zero(coeffs.back())) coeffs.pop_back();
    return coeffs;
}

// Coefficients of the unique polynomial of degree <= d through (0,y0)..(d,yd), trailing zeros
// stripped ([ARK] DensePolynomial::from_coefficients_vec).  The polynomial is unique, so this equals
// the reference's IFFT over the size-4 domain (triangle-counting/src/lib.rs:121-131,
// gkr-protocol/src/round_polynomial.rs:79-89) coefficient for coefficient.
inline std::vector<Fe> lagrange_to_coeffs(const HostField& F, const std::vector<Fe>& ys) {
    const size_t n = ys.size();
    std::vector<Fe> coeffs(n, F.zero());
    for (size_t i = 0; i < n; ++i) {
        std::vector<Fe> num(1, F.one());
        Fe den = F.one();
        for (size_t j = 0; j < n; ++j) {
            if (j == i) continue;
            Fe fj = F.from_u64(j);
            std::vector<Fe> nn(num.size() + 1, F.zero());
            for (size_t k = 0; k < num.size(); ++k) {
                nn[k + 1] = F.add(nn[k + 1], num[k]);
                nn[k] = F.sub(nn[k], F.mul(fj, num[k]));
            }
            num.swap(nn);
            den = F.mul(den, F.sub(F.from_u64(i), fj));
        }
        Fe s = F.mul(ys[i], F.inverse(den));
        for (size_t k = 0; k < n; ++k) coeffs[k] = F.add(coeffs[k], F.mul(s, num[k]));
    }
    while (!coeffs.empty() && F.is_zero(coeffs.back())) coeffs.pop_back();
    return coeffs;
}

}  // namespace scb

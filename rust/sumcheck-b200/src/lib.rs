//! SOURCE ONLY -- not compiled in this environment (no Rust toolchain).
//!
//! `GpuPoly<F>`: an implementor of the reference's `sum_check_protocol::SumCheckPolynomial<F>`
//! (sum-check-protocol/src/lib.rs:121-156) whose tables live in B200 HBM.  With it the reference's
//! `Prover`, `Verifier` (:73-117, :227-331) and `fiat_shamir::generate_transcript` run unchanged:
//!
//! ```ignore
//! let g = GpuPoly::<Fp389>::product(&[&a, &b, &c]);      // or GpuPoly::matmul_g(n, a_iter, b_iter, &point)
//! let mut prover = sum_check_protocol::Prover::new(g.clone());
//! let transcript = fiat_shamir::generate_transcript::<_, _, DefaultFieldHasher<Sha256>>(prover)?;
//! ```
//!
//! Data crosses the FFI in ark-ff's in-memory format: `Fp<MontBackend<C,N>,N>` is `#[repr(transparent)]`-like
//! over `BigInt<N>([u64; N])` holding the Montgomery form, so `&[F]` is passed as `*const u64`.
use ark_ff::{BigInteger, PrimeField};
use ark_poly::{univariate::SparsePolynomial, DenseUVPolynomial};
use std::{marker::PhantomData, os::raw::c_int, ptr};
use sum_check_protocol::SumCheckPolynomial;

#[repr(C)] pub struct scb_field { _p: [u8; 0] }
#[repr(C)] pub struct scb_mle { _p: [u8; 0] }
#[repr(C)] pub struct scb_poly { _p: [u8; 0] }

// The subset of include/sumcheck_b200.h the trait impl needs.
extern "C" {
    fn scb_field_create(n_limbs: u32, modulus_le: *const u64, out: *mut *mut scb_field) -> c_int;
    fn scb_field_free(f: *mut scb_field);
    fn scb_mle_from_host(f: *const scb_field, num_vars: u32, evals: *const u64, out: *mut *mut scb_mle) -> c_int;
    fn scb_mle_free(m: *mut scb_mle);
    fn scb_poly_product(tables: *const *const scb_mle, k: u32, out: *mut *mut scb_poly) -> c_int;
    fn scb_poly_product_from_host(f: *const scb_field, k: u32, num_vars: u32, host_tables: *const *const u64, out: *mut *mut scb_poly) -> c_int;
    fn scb_poly_matmul_g_new(f: *const scb_field, n: u32, a: *const u64, b: *const u64, point: *const u64,
                             out: *mut *mut scb_poly) -> c_int;
    fn scb_poly_clone(p: *const scb_poly, out: *mut *mut scb_poly) -> c_int;
    fn scb_poly_free(p: *mut scb_poly);
    fn scb_poly_num_vars(p: *const scb_poly, out: *mut u32) -> c_int;
    fn scb_poly_evaluate(p: *const scb_poly, point: *const u64, n: u32, out: *mut u64) -> c_int;
    fn scb_poly_fix_variables(p: *const scb_poly, pp: *const u64, n: u32, out: *mut *mut scb_poly) -> c_int;
    fn scb_poly_to_evaluations(p: *const scb_poly, out: *mut u64, cap: usize) -> c_int;
    fn scb_poly_to_univariate(p: *const scb_poly, degrees: *mut u64, coeffs: *mut u64, cap: u32, n: *mut u32) -> c_int;
    fn scb_last_error() -> *const std::os::raw::c_char;
}

const SCB_EINVAL: c_int = -1;

fn check(rc: c_int) {
    if rc != 0 {
        let msg = unsafe { std::ffi::CStr::from_ptr(scb_last_error()) }.to_string_lossy().into_owned();
        panic!("sumcheck_b200: [{rc}] {msg}"); // the reference's provers panic via unwrap/assert (gkr-protocol/src/lib.rs:127,157)
    }
}

fn as_words<F: PrimeField>(s: &[F]) -> *const u64 {
    // Fp<MontBackend<_,N>,N> = BigInt<N>([u64; N]) in Montgomery form (+ zero-sized PhantomData)
    debug_assert_eq!(std::mem::size_of::<F>(), 8 * F::BigInt::NUM_LIMBS);
    s.as_ptr() as *const u64
}

pub struct GpuPoly<F: PrimeField> {
    field: *mut scb_field,
    poly: *mut scb_poly,
    owns_field: bool,
    _f: PhantomData<F>,
}

impl<F: PrimeField> GpuPoly<F> {
    fn field() -> *mut scb_field {
        let modulus = F::MODULUS;
        let limbs: &[u64] = modulus.as_ref();
        let mut f = ptr::null_mut();
        check(unsafe { scb_field_create(limbs.len() as u32, limbs.as_ptr(), &mut f) });
        f
    }

    /// Product of K dense multilinear tables over the same variables (`ProductMLE<K>`): one call takes all K slices
    /// as they lie in memory; small-prime tables are narrowed on their way to the device (include/sumcheck_b200.h).
    pub fn product(tables: &[&[F]]) -> Self {
        let field = Self::field();
        let nv = tables[0].len().trailing_zeros();
        assert!(tables.iter().all(|t| t.len() == 1usize << nv), "The size of evaluations should be 2^num_vars.");
        let ptrs: Vec<*const u64> = tables.iter().map(|t| as_words(t)).collect();
        let mut poly = ptr::null_mut();
        check(unsafe { scb_poly_product_from_host(field, ptrs.len() as u32, nv, ptrs.as_ptr(), &mut poly) });
        Self { field, poly, owns_field: true, _f: PhantomData }
    }

    /// The same from tables that are already on the device (`scb_mle` handles are built one by one).
    pub fn product_of_mles(tables: &[&[F]]) -> Self {
        let field = Self::field();
        let nv = tables[0].len().trailing_zeros();
        let mut mles = vec![];
        for t in tables {
            let mut m = ptr::null_mut();
            check(unsafe { scb_mle_from_host(field, nv, as_words(t), &mut m) });
            mles.push(m as *const scb_mle);
        }
        let mut poly = ptr::null_mut();
        check(unsafe { scb_poly_product(mles.as_ptr(), mles.len() as u32, &mut poly) });
        for m in mles { unsafe { scb_mle_free(m as *mut scb_mle) }; }
        Self { field, poly, owns_field: true, _f: PhantomData }
    }

    /// `matrix_multiplication::G::new(n, a, b, point)` (matrix-multiplication/src/lib.rs:77-92).
    pub fn matmul_g(n: usize, a: impl IntoIterator<Item = F>, b: impl IntoIterator<Item = F>, point: &[F]) -> Self {
        let field = Self::field();
        let (a, b): (Vec<F>, Vec<F>) = (a.into_iter().collect(), b.into_iter().collect());
        let mut poly = ptr::null_mut();
        check(unsafe { scb_poly_matmul_g_new(field, n as u32, as_words(&a), as_words(&b), as_words(point), &mut poly) });
        Self { field, poly, owns_field: true, _f: PhantomData }
    }
}

impl<F: PrimeField> Clone for GpuPoly<F> {
    fn clone(&self) -> Self {
        let mut poly = ptr::null_mut();
        check(unsafe { scb_poly_clone(self.poly, &mut poly) }); // O(1): shares the device tables
        Self { field: self.field, poly, owns_field: false, _f: PhantomData }
    }
}

impl<F: PrimeField> Drop for GpuPoly<F> {
    fn drop(&mut self) {
        unsafe {
            scb_poly_free(self.poly);
            if self.owns_field { scb_field_free(self.field) } // handles keep the field alive internally
        }
    }
}

impl<F: PrimeField> SumCheckPolynomial<F> for GpuPoly<F> {
    fn evaluate(&self, point: &[F]) -> Option<F> {
        let mut out = F::zero();
        match unsafe { scb_poly_evaluate(self.poly, as_words(point), point.len() as u32, &mut out as *mut F as *mut u64) } {
            0 => Some(out),
            SCB_EINVAL => None, // dimension mismatch (:124-126)
            rc => { check(rc); None }
        }
    }

    fn fix_variables(&self, partial_point: &[F]) -> Self {
        let mut poly = ptr::null_mut();
        check(unsafe { scb_poly_fix_variables(self.poly, as_words(partial_point), partial_point.len() as u32, &mut poly) });
        Self { field: self.field, poly, owns_field: false, _f: PhantomData }
    }

    fn to_univariate(&self) -> SparsePolynomial<F> {
        let (mut deg, mut co, mut n) = ([0u64; 8], vec![F::zero(); 8], 0u32);
        check(unsafe { scb_poly_to_univariate(self.poly, deg.as_mut_ptr(), co.as_mut_ptr() as *mut u64, 8, &mut n) });
        // The library already applied the implementor's zero-term conventions; rebuild the term list verbatim.
        let terms: Vec<(usize, F)> = (0..n as usize).map(|i| (deg[i] as usize, co[i])).collect();
        // NOTE: from_coefficients_vec would pop a trailing explicit zero; the reference's own G can never end
        // in one (its last term is the non-zero x^2 coefficient or the list is empty), so this is exact.
        SparsePolynomial::from_coefficients_vec(terms)
    }

    fn num_vars(&self) -> usize {
        let mut n = 0u32;
        check(unsafe { scb_poly_num_vars(self.poly, &mut n) });
        n as usize
    }

    fn to_evaluations(&self) -> Vec<F> {
        let len = 1usize << self.num_vars();
        let mut v = vec![F::zero(); len];
        check(unsafe { scb_poly_to_evaluations(self.poly, v.as_mut_ptr() as *mut u64, len) });
        v
    }
}

// =====================================================================================================================
// The FAST path: replacements for `sum_check_protocol::Prover::{new, round}` (sum-check-protocol/src/lib.rs:88-112)
// and `fiat_shamir::generate_transcript` (fiat-shamir/src/lib.rs:75-98) on top of the library's own prover object.
// `Prover<F, GpuPoly<F>>` above is the drop-in through the trait only: it pays `to_evaluations()` (a 2^v-entry D2H) in
// `Prover::new` and two launches per round.  `GpuProver<F>` keeps everything on the device: c_1 comes from the first
// pass (which also yields g_1, and g_2 for small-prime fields), every later round is ONE fused fold+message pass, and
// under Fiat-Shamir (`generate_transcript_gpu`) the whole proof runs in resident kernels.  Same messages, same bytes
// (bench.py's `e2e.trait_only.equals_fast_path_bytes`, tests/test_gpu_trait_path.py).
// =====================================================================================================================
#[repr(C)] pub struct scb_prover { _p: [u8; 0] }

/// `scb_pair_cb` of include/sumcheck_b200.h
pub type ScbPairCb = unsafe extern "C" fn(user: *mut std::os::raw::c_void, pass: u32, n_vals: u32, vals: *const u64, next_pair_out: *mut u64) -> c_int;

extern "C" {
    fn scb_prover_new(g: *const scb_poly, out: *mut *mut scb_prover) -> c_int;                       // Prover::new  :88-97
    fn scb_prover_free(p: *mut scb_prover);
    fn scb_prover_c_1(p: *const scb_prover, out: *mut u64) -> c_int;                                 // :100-102
    fn scb_prover_num_vars(p: *const scb_prover, out: *mut u32) -> c_int;                            // :114-116
    fn scb_prover_round(p: *mut scb_prover, r_prev: *const u64, j: u32, degrees: *mut u64, coeffs: *mut u64, cap: u32, n: *mut u32) -> c_int; // :105-112
    fn scb_fs_generate_transcript(p: *mut scb_prover, out: *mut u8, cap: usize, out_len: *mut usize, offsets: *mut u64) -> c_int; // fiat-shamir :75-98
    fn scb_poly_n_points(p: *const scb_poly, out: *mut u32) -> c_int;
    fn scb_poly_grid_evals(p: *const scb_poly, out: *mut u64) -> c_int;
    fn scb_poly_resident_pairs(p: *const scb_poly, ra: *const u64, rb: *const u64, max_passes: u32, cb: ScbPairCb, user: *mut std::os::raw::c_void,
                               passes_done: *mut u32, out_folded: *mut *mut scb_poly) -> c_int;
    fn scb_poly_fix_and_round_evals_claim(p: *const scb_poly, r: *const u64, claim: *const u64, n_points: u32, out: *mut *mut scb_poly, out_elems: *mut u64) -> c_int;
    fn scb_set_option(name: *const std::os::raw::c_char, value: i64) -> c_int;
}

/// Library switches (csrc/options.hpp).  The library never reads the environment; a host that wants a switch sets it here.
pub fn set_option(name: &str, value: i64) {
    let c = std::ffi::CString::new(name).unwrap();
    check(unsafe { scb_set_option(c.as_ptr(), value) });
}

/// Drop-in for `sum_check_protocol::Prover<F, P>` with the same four methods.
pub struct GpuProver<F: PrimeField> {
    prover: *mut scb_prover,
    _g: GpuPoly<F>, // keeps the field handle alive
}

impl<F: PrimeField> GpuProver<F> {
    /// `Prover::new(g)`: c_1 without the 2^v-entry host `Vec` (:89).
    pub fn new(g: GpuPoly<F>) -> Self {
        let mut prover = ptr::null_mut();
        check(unsafe { scb_prover_new(g.poly, &mut prover) });
        Self { prover, _g: g }
    }
    pub fn c_1(&self) -> F {
        let mut out = F::zero();
        check(unsafe { scb_prover_c_1(self.prover, &mut out as *mut F as *mut u64) });
        out
    }
    /// `Prover::round(r_prev, j)`: fold + message in one pass for j >= 1.
    pub fn round(&mut self, r_prev: F, j: usize) -> SparsePolynomial<F> {
        let (mut deg, mut co, mut n) = ([0u64; 8], vec![F::zero(); 8], 0u32);
        check(unsafe { scb_prover_round(self.prover, &r_prev as *const F as *const u64, j as u32, deg.as_mut_ptr(), co.as_mut_ptr() as *mut u64, 8, &mut n) });
        SparsePolynomial::from_coefficients_vec((0..n as usize).map(|i| (deg[i] as usize, co[i])).collect())
    }
    pub fn num_vars(&self) -> usize {
        let mut n = 0u32;
        check(unsafe { scb_prover_num_vars(self.prover, &mut n) });
        n as usize
    }
}
impl<F: PrimeField> Drop for GpuProver<F> {
    fn drop(&mut self) { unsafe { scb_prover_free(self.prover) } }
}

/// With this impl the reference's own `fiat_shamir::generate_transcript::<F, GpuProver<F>, H>` runs unchanged, for any
/// hasher H, one fused pass per round.  (Add `fiat-shamir`, `ark-serialize` to [dependencies].)
impl<F: PrimeField> fiat_shamir::InteractiveProver<F> for GpuProver<F> {
    fn g_1(&mut self) -> fiat_shamir::Result<Vec<u8>> {
        use ark_serialize::CanonicalSerialize;
        let mut res = vec![];
        let p: (F, SparsePolynomial<F>) = (self.c_1(), GpuProver::round(self, F::one(), 0));
        p.serialize_uncompressed(&mut res)?;
        Ok(res)
    }
    fn round(&mut self, j: usize, r_j: F) -> fiat_shamir::Result<Vec<u8>> {
        use ark_serialize::CanonicalSerialize;
        let mut res = vec![];
        GpuProver::round(self, r_j, j).serialize_uncompressed(&mut res)?;
        Ok(res)
    }
    fn num_rounds(&self) -> usize { self.num_vars() }
}

/// `fiat_shamir::generate_transcript::<F, Prover<F, P>, DefaultFieldHasher<Sha256>>` (fiat-shamir/src/lib.rs:75-98) with
/// the hash chain inside the library: resident kernels, two rounds per pass for small-prime fields, no per-round launch
/// or stream synchronisation.  Returns the messages g_1, g_2, ... exactly as `FiatShamirTranscript.g` holds them.
pub fn generate_transcript_gpu<F: PrimeField>(prover: GpuProver<F>) -> Vec<Vec<u8>> {
    let nv = prover.num_vars();
    let elem = (F::MODULUS_BIT_SIZE as usize + 7) / 8;
    let cap = 64 + nv * (8 + 8 * (8 + elem)) + elem;
    let (mut buf, mut len, mut offs) = (vec![0u8; cap], 0usize, vec![0u64; nv + 2]);
    check(unsafe { scb_fs_generate_transcript(prover.prover, buf.as_mut_ptr(), cap, &mut len, offs.as_mut_ptr()) });
    (0..nv.max(1)).map(|i| buf[offs[i] as usize..offs[i + 1] as usize].to_vec()).collect()
}

/// The callback form for a host that owns the transcript (any hash, any message format): Prover::new's grid, then every
/// pass of the proof inside ONE resident kernel; `next` receives a pass's (K+1)^2 grid sums H[a][b] (or the K+1 line sums
/// when a single variable is left) and returns the next two challenges.  Small-prime fields, product polynomials
/// (include/sumcheck_b200.h: scb_poly_grid_evals / scb_poly_resident_pairs; csrc/pairs.cuh for the algebra:
/// g_j(X) = H(X,0) + H(X,1), g_{j+1}(Y) = H(r_j, Y)).
pub fn prove_with_pair_callback<F: PrimeField, N: FnMut(u32, &[F]) -> (F, F)>(g: &GpuPoly<F>, first_pair: impl FnOnce(&[F]) -> (F, F), mut next: N) {
    unsafe extern "C" fn tramp<F: PrimeField, N: FnMut(u32, &[F]) -> (F, F)>(user: *mut std::os::raw::c_void, pass: u32, n_vals: u32, vals: *const u64,
                                                                            out: *mut u64) -> c_int {
        let f = &mut *(user as *mut N);
        let vals = std::slice::from_raw_parts(vals as *const F, n_vals as usize);
        let (ra, rb) = f(pass, vals);
        *(out as *mut F) = ra;
        *(out as *mut F).add(1) = rb;
        0
    }
    let mut np = 0u32;
    check(unsafe { scb_poly_n_points(g.poly, &mut np) });
    let mut grid = vec![F::zero(); (np * np) as usize];
    check(unsafe { scb_poly_grid_evals(g.poly, grid.as_mut_ptr() as *mut u64) });
    let (ra, rb) = first_pair(&grid);
    let mut done = 0u32;
    check(unsafe {
        scb_poly_resident_pairs(g.poly, &ra as *const F as *const u64, &rb as *const F as *const u64, 0, tramp::<F, N>,
                                &mut next as *mut N as *mut std::os::raw::c_void, &mut done, ptr::null_mut())
    });
}

/// One fused round for a host-driven prover that knows the claim g_j(0) + g_j(1) = g_{j-1}(r_{j-1}): 4-limb fields skip a
/// point (csrc/g4.cuh).  Returns (folded polynomial, sums at X = 0..d).
pub fn fix_and_round_with_claim<F: PrimeField>(g: &GpuPoly<F>, r: F, claim: F) -> (GpuPoly<F>, Vec<F>) {
    let mut np = 0u32;
    check(unsafe { scb_poly_n_points(g.poly, &mut np) });
    let mut out = vec![F::zero(); np as usize];
    let mut poly = ptr::null_mut();
    check(unsafe { scb_poly_fix_and_round_evals_claim(g.poly, &r as *const F as *const u64, &claim as *const F as *const u64, np, &mut poly, out.as_mut_ptr() as *mut u64) });
    (GpuPoly { field: g.field, poly, owns_field: false, _f: PhantomData }, out)
}

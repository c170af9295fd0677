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

// N-API shim over the C ABI (include/bls381_b200.h): every native the TypeScript wrapper js/index.ts needs.
//
// This image has no node and no node_api.h, so the addon cannot be built or run here; what CAN be checked
// is checked: tests/test_js_shim.py compiles this file with `g++ -fsyntax-only` against js/stub/node_api.h (declarations of
// exactly the N-API functions used below, signatures as in Node's node_api.h / js_native_api.h) and checks that the set of
// exported names equals the set of `native.*` calls in js/index.ts.  The binding exercised by the test-suite is the ctypes
// one (noble_bls12_381_b200/_lib.py).  Build on a machine with node >= 18:  cd js && npm install && npx node-gyp rebuild
//
// Conventions: every buffer length and typed-array type is validated BEFORE the native call (RangeError / TypeError): the C
// ABI trusts its sizes.  Synchronous natives call straight through; the Promise-returning functions of the reference (sign /
// verify / verifyBatch, index.ts:746, 756, 792) run on the libuv pool (napi_create_async_work) with COPIES of their inputs
// (the reference never retains caller buffers either, index.ts:159-163), so the JS thread never blocks on a CUDA stream.
// BLS381_B200_DEVICES=<bit mask> selects the GPUs (default: device 0); with more than one, verifyBatch shards across them
// (bls381_verify_batch_multi: one all-gather of the partial products over NVLink).
#include <node_api.h>

#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "bls381_b200.h"

namespace {

struct Bytes { const uint8_t* p; size_t n; };

bool get_bytes(napi_env env, napi_value v, Bytes* out) {
  bool is_ta = false;
  if (napi_is_typedarray(env, v, &is_ta) != napi_ok || !is_ta) { napi_throw_type_error(env, nullptr, "expected Uint8Array"); return false; }
  napi_typedarray_type t; napi_value ab; size_t off; void* data = nullptr;
  napi_get_typedarray_info(env, v, &t, &out->n, &data, &ab, &off);
  if (t != napi_uint8_array) { napi_throw_type_error(env, nullptr, "expected Uint8Array"); return false; }
  out->p = static_cast<const uint8_t*>(data);
  return true;
}

// BigUint64Array of n + 1 byte offsets: off[0] == 0, non-decreasing, off[n] == total
bool get_offsets(napi_env env, napi_value v, size_t total, std::vector<uint64_t>* out) {
  bool is_ta = false;
  if (napi_is_typedarray(env, v, &is_ta) != napi_ok || !is_ta) { napi_throw_type_error(env, nullptr, "expected BigUint64Array"); return false; }
  napi_typedarray_type t; napi_value ab; size_t off, len; void* data = nullptr;
  napi_get_typedarray_info(env, v, &t, &len, &data, &ab, &off);
  if (t != napi_biguint64_array) { napi_throw_type_error(env, nullptr, "expected BigUint64Array"); return false; }
  if (len < 1) { napi_throw_range_error(env, nullptr, "offsets: need n + 1 entries"); return false; }
  const uint64_t* p = static_cast<const uint64_t*>(data);
  if (p[0] != 0 || p[len - 1] != total) { napi_throw_range_error(env, nullptr, "offsets do not span the message buffer"); return false; }
  for (size_t i = 1; i < len; ++i)
    if (p[i] < p[i - 1]) { napi_throw_range_error(env, nullptr, "offsets must be non-decreasing"); return false; }
  out->assign(p, p + len);
  return true;
}

napi_value make_u8(napi_env env, size_t n, uint8_t** data) {
  napi_value ab, ta;
  void* p = nullptr;
  napi_create_arraybuffer(env, n, &p, &ab);
  napi_create_typedarray(env, napi_uint8_array, n, ab, 0, &ta);
  *data = static_cast<uint8_t*>(p);
  return ta;
}

napi_value make_i32(napi_env env, size_t n, int32_t** data) {
  napi_value ab, ta;
  void* p = nullptr;
  napi_create_arraybuffer(env, n * 4, &p, &ab);
  napi_create_typedarray(env, napi_int32_array, n, ab, 0, &ta);
  *data = static_cast<int32_t*>(p);
  return ta;
}

napi_value fail(napi_env env) { napi_throw_error(env, nullptr, bls381_last_error()); return nullptr; }
napi_value range(napi_env env, const char* msg) { napi_throw_range_error(env, nullptr, msg); return nullptr; }

napi_value pair_obj(napi_env env, const char* k1, napi_value v1, const char* k2, napi_value v2) {
  napi_value obj;
  napi_create_object(env, &obj);
  napi_set_named_property(env, obj, k1, v1);
  napi_set_named_property(env, obj, k2, v2);
  return obj;
}

// pairingBatch(g1: Uint8Array(n*96), g2: Uint8Array(n*192), withFinalExponent: boolean, checked: boolean)
//   -> { out: Uint8Array(n*576), status: Int32Array(n) }      (status all 0 when checked == false)
napi_value PairingBatch(napi_env env, napi_callback_info info) {
  size_t argc = 4; napi_value a[4];
  napi_get_cb_info(env, info, &argc, a, nullptr, nullptr);
  if (argc < 2) return range(env, "pairingBatch(g1, g2, withFinalExponent, checked)");
  Bytes g1, g2; bool fe = true, checked = true;
  if (!get_bytes(env, a[0], &g1) || !get_bytes(env, a[1], &g2)) return nullptr;
  if (argc > 2) napi_get_value_bool(env, a[2], &fe);
  if (argc > 3) napi_get_value_bool(env, a[3], &checked);
  const size_t n = g1.n / 96;
  if (n == 0 || g1.n != n * 96 || g2.n != n * 192) return range(env, "pairingBatch: g1 must be n x 96 bytes and g2 n x 192 bytes");
  uint8_t* out; napi_value r = make_u8(env, n * 576, &out);
  int32_t* st; napi_value s = make_i32(env, n, &st);
  memset(st, 0, n * 4);
  if (bls381_pairing_batch(g1.p, g2.p, n, fe, out, checked ? st : nullptr) != 0) return fail(env);
  return pair_obj(env, "out", r, "status", s);
}

// millerProduct(g1, g2, withFinalExponent) -> Uint8Array(576)     (core of verify / verifyBatch, index.ts:763-766, 812-817)
napi_value MillerProduct(napi_env env, napi_callback_info info) {
  size_t argc = 3; napi_value a[3];
  napi_get_cb_info(env, info, &argc, a, nullptr, nullptr);
  if (argc < 2) return range(env, "millerProduct(g1, g2, withFinalExponent)");
  Bytes g1, g2; bool fe = true;
  if (!get_bytes(env, a[0], &g1) || !get_bytes(env, a[1], &g2)) return nullptr;
  if (argc > 2) napi_get_value_bool(env, a[2], &fe);
  const size_t n = g1.n / 96;
  if (n == 0 || g1.n != n * 96 || g2.n != n * 192) return range(env, "millerProduct: g1 must be n x 96 bytes and g2 n x 192 bytes");
  uint8_t* out; napi_value r = make_u8(env, 576, &out);
  if (bls381_miller_product(g1.p, g2.p, n, fe, out) != 0) return fail(env);
  return r;
}

// finalExpBatch(f12: Uint8Array(n*576)) -> Uint8Array(n*576)      (Fp12#finalExponentiate, math.ts:856-874)
napi_value FinalExpBatch(napi_env env, napi_callback_info info) {
  size_t argc = 1; napi_value a[1];
  napi_get_cb_info(env, info, &argc, a, nullptr, nullptr);
  Bytes in;
  if (argc < 1 || !get_bytes(env, a[0], &in)) return nullptr;
  const size_t n = in.n / 576;
  if (n == 0 || in.n != n * 576) return range(env, "finalExpBatch: expected n x 576 bytes");
  uint8_t* out; napi_value r = make_u8(env, n * 576, &out);
  if (bls381_final_exp_batch(in.p, n, out) != 0) return fail(env);
  return r;
}

// fp12Product(f12: Uint8Array(n*576), withFinalExponent) -> Uint8Array(576)
napi_value Fp12Product(napi_env env, napi_callback_info info) {
  size_t argc = 2; napi_value a[2];
  napi_get_cb_info(env, info, &argc, a, nullptr, nullptr);
  Bytes in; bool fe = false;
  if (argc < 1 || !get_bytes(env, a[0], &in)) return nullptr;
  if (argc > 1) napi_get_value_bool(env, a[1], &fe);
  const size_t n = in.n / 576;
  if (n == 0 || in.n != n * 576) return range(env, "fp12Product: expected n x 576 bytes");
  uint8_t* out; napi_value r = make_u8(env, 576, &out);
  if (bls381_fp12_product(in.p, n, fe, out) != 0) return fail(env);
  return r;
}

// g1Decompress(keys: Uint8Array(n*48)) -> { points: Uint8Array(n*96), status: Int32Array(n) }   (PointG1.fromHex, index.ts:298-327)
// g2Decompress(sigs: Uint8Array(n*96)) -> { points: Uint8Array(n*192), status: Int32Array(n) }  (PointG2.fromSignature, :500-530)
template <int IN, int OUT, int (*FN)(const uint8_t*, size_t, uint8_t*, int32_t*)>
napi_value Decompress(napi_env env, napi_callback_info info) {
  size_t argc = 1; napi_value a[1];
  napi_get_cb_info(env, info, &argc, a, nullptr, nullptr);
  Bytes in;
  if (argc < 1 || !get_bytes(env, a[0], &in)) return nullptr;
  const size_t n = in.n / IN;
  if (n == 0 || in.n != n * IN) return range(env, "decompress: bad input length");
  uint8_t* out; napi_value pts = make_u8(env, n * OUT, &out);
  int32_t* sp; napi_value st = make_i32(env, n, &sp);
  if (FN(in.p, n, out, sp) != 0) return fail(env);
  return pair_obj(env, "points", pts, "status", st);
}

// g1Validate(points: Uint8Array(n*96)) / g2Validate(points: Uint8Array(n*192)) -> Int32Array(n)   (assertValidity, :383-388 / :633-638)
template <int IN, int (*FN)(const uint8_t*, size_t, int32_t*)>
napi_value Validate(napi_env env, napi_callback_info info) {
  size_t argc = 1; napi_value a[1];
  napi_get_cb_info(env, info, &argc, a, nullptr, nullptr);
  Bytes in;
  if (argc < 1 || !get_bytes(env, a[0], &in)) return nullptr;
  const size_t n = in.n / IN;
  if (n == 0 || in.n != n * IN) return range(env, "validate: bad input length");
  int32_t* sp; napi_value st = make_i32(env, n, &sp);
  if (FN(in.p, n, sp) != 0) return fail(env);
  return st;
}

// aggregateG1(keys: Uint8Array(n*48)) -> { point: Uint8Array(48), status: Int32Array(n) }    (aggregatePublicKeys, :773-778)
// aggregateG2(sigs: Uint8Array(n*96)) -> { point: Uint8Array(96), status: Int32Array(n) }    (aggregateSignatures, :783-788)
template <int IN, int (*FN)(const uint8_t*, size_t, uint8_t*, int32_t*)>
napi_value Aggregate(napi_env env, napi_callback_info info) {
  size_t argc = 1; napi_value a[1];
  napi_get_cb_info(env, info, &argc, a, nullptr, nullptr);
  Bytes in;
  if (argc < 1 || !get_bytes(env, a[0], &in)) return nullptr;
  const size_t n = in.n / IN;
  if (n == 0 || in.n != n * IN) return range(env, "aggregate: bad input length");
  uint8_t* out; napi_value pt = make_u8(env, IN, &out);
  int32_t* sp; napi_value st = make_i32(env, n, &sp);
  if (FN(in.p, n, out, sp) != 0) return fail(env);
  return pair_obj(env, "point", pt, "status", st);
}

// g1ScalarMul(points: Uint8Array(n*96), scalars: Uint8Array(n*32)) -> { points, flags }   (PointG1#multiply, math.ts:1061-1078)
// g2ScalarMul(points: Uint8Array(n*192), scalars: Uint8Array(n*32)) -> { points, flags }  (PointG2#multiply; sign(PointG2, key))
template <int PT, int (*FN)(const uint8_t*, const uint8_t*, size_t, uint8_t*, int32_t*)>
napi_value ScalarMul(napi_env env, napi_callback_info info) {
  size_t argc = 2; napi_value a[2];
  napi_get_cb_info(env, info, &argc, a, nullptr, nullptr);
  Bytes pts, sc;
  if (argc < 2 || !get_bytes(env, a[0], &pts) || !get_bytes(env, a[1], &sc)) return nullptr;
  const size_t n = sc.n / 32;
  if (n == 0 || sc.n != n * 32 || pts.n != n * PT) return range(env, "scalarMul: bad input lengths");
  uint8_t* out; napi_value r = make_u8(env, n * PT, &out);
  int32_t* fp; napi_value fl = make_i32(env, n, &fp);
  if (FN(pts.p, sc.p, n, out, fp) != 0) return fail(env);
  return pair_obj(env, "points", r, "flags", fl);
}

// getPublicKeyBatch(sks: Uint8Array(n*32)) -> Uint8Array(n*48)       (getPublicKey, index.ts:738-740)
napi_value GetPublicKeyBatch(napi_env env, napi_callback_info info) {
  size_t argc = 1; napi_value a[1];
  napi_get_cb_info(env, info, &argc, a, nullptr, nullptr);
  Bytes sk;
  if (argc < 1 || !get_bytes(env, a[0], &sk)) return nullptr;
  const size_t n = sk.n / 32;
  if (n == 0 || sk.n != n * 32) return range(env, "getPublicKeyBatch: expected n x 32 bytes");
  uint8_t* out; napi_value r = make_u8(env, n * 48, &out);
  if (bls381_get_public_key_batch(sk.p, n, out) != 0) return fail(env);
  return r;
}

// hashToG2(packedMsgs: Uint8Array, offsets: BigUint64Array(n+1), dst: Uint8Array) -> Uint8Array(n*192)  (PointG2.hashToCurve, :481-490)
// hashToG1(packedMsgs, offsets, dst) -> Uint8Array(n*96)                                                 (PointG1.hashToCurve, :331-339)
template <int OUT, int (*FN)(const uint8_t*, const uint64_t*, size_t, const uint8_t*, size_t, uint8_t*)>
napi_value HashToCurve(napi_env env, napi_callback_info info) {
  size_t argc = 3; napi_value a[3];
  napi_get_cb_info(env, info, &argc, a, nullptr, nullptr);
  Bytes msgs, dst; std::vector<uint64_t> off;
  if (argc < 3 || !get_bytes(env, a[0], &msgs) || !get_offsets(env, a[1], msgs.n, &off) || !get_bytes(env, a[2], &dst)) return nullptr;
  const size_t n = off.size() - 1;
  if (n == 0) return range(env, "hashToCurve: empty batch");
  uint8_t* out; napi_value r = make_u8(env, n * OUT, &out);
  if (FN(msgs.p, off.data(), n, dst.p, dst.n, out) != 0) return fail(env);
  return r;
}

// g1Encode(points: Uint8Array(n*96), compressed: boolean) -> Uint8Array(n*48 | n*96)     (PointG1#toRawBytes, :355-376)
// g2Encode(points: Uint8Array(n*192), compressed: boolean) -> Uint8Array(n*96 | n*192)   (PointG2#toSignature / toRawBytes, :586-631)
template <int PT, int (*FN)(const uint8_t*, size_t, int, uint8_t*)>
napi_value Encode(napi_env env, napi_callback_info info) {
  size_t argc = 2; napi_value a[2];
  napi_get_cb_info(env, info, &argc, a, nullptr, nullptr);
  Bytes pts; bool compressed = false;
  if (argc < 2 || !get_bytes(env, a[0], &pts)) return nullptr;
  if (napi_get_value_bool(env, a[1], &compressed) != napi_ok) { napi_throw_type_error(env, nullptr, "encode: boolean expected"); return nullptr; }
  const size_t n = pts.n / PT;
  if (n == 0 || pts.n != n * PT) return range(env, "encode: bad input length");
  uint8_t* out; napi_value r = make_u8(env, n * (compressed ? PT / 2 : PT), &out);
  if (FN(pts.p, n, compressed ? 1 : 0, out) != 0) return fail(env);
  return r;
}

// ---- async work -------------------------------------------------------------------------------------------------
// verifyBatch(sig: Uint8Array(96), packedMsgs, offsets: BigUint64Array(n+1), pks: Uint8Array(n*48), dst) -> Promise<{verdict, status}>
// signBatch(sks: Uint8Array(n*32), packedMsgs, offsets, dst) -> Promise<Uint8Array(n*96)>
struct Job {
  napi_async_work work = nullptr; napi_deferred deferred = nullptr;
  bool is_sign = false;
  std::vector<uint8_t> a, msgs, pks, dst, out; std::vector<uint64_t> off; std::vector<int32_t> status;
  int verdict = 0, rc = 0; std::string err;
};
void JobExec(napi_env, void* d) {
  auto* j = static_cast<Job*>(d);
  const size_t n = j->off.size() - 1;
  if (j->is_sign) {
    j->out.resize(n * 96);
    j->rc = bls381_sign_batch(j->a.data(), j->msgs.data(), j->off.data(), n, j->dst.data(), j->dst.size(), j->out.data());
    if (!j->a.empty()) memset(j->a.data(), 0, j->a.size());  // the private keys
  } else {
    j->status.resize(n + 1);
    auto fn = bls381_device_count() > 1 ? bls381_verify_batch_multi : bls381_verify_batch;
    j->rc = fn(j->a.data(), j->msgs.data(), j->off.data(), j->pks.data(), n, j->dst.data(), j->dst.size(), &j->verdict, j->status.data());
  }
  if (j->rc) j->err = bls381_last_error();
}
void JobDone(napi_env env, napi_status, void* d) {
  auto* j = static_cast<Job*>(d);
  if (j->rc) {
    napi_value msg, e;
    napi_create_string_utf8(env, j->err.c_str(), NAPI_AUTO_LENGTH, &msg);
    napi_create_error(env, nullptr, msg, &e);
    napi_reject_deferred(env, j->deferred, e);
  } else if (j->is_sign) {
    uint8_t* p; napi_value r = make_u8(env, j->out.size(), &p);
    memcpy(p, j->out.data(), j->out.size());
    napi_resolve_deferred(env, j->deferred, r);
  } else {
    napi_value v; int32_t* sp;
    napi_create_int32(env, j->verdict, &v);
    napi_value st = make_i32(env, j->status.size(), &sp);
    memcpy(sp, j->status.data(), j->status.size() * 4);
    napi_resolve_deferred(env, j->deferred, pair_obj(env, "verdict", v, "status", st));
  }
  napi_delete_async_work(env, j->work);
  delete j;
}
napi_value queue(napi_env env, Job* j, const char* name_) {
  napi_value promise, name;
  napi_create_promise(env, &j->deferred, &promise);
  napi_create_string_utf8(env, name_, NAPI_AUTO_LENGTH, &name);
  napi_create_async_work(env, nullptr, name, JobExec, JobDone, j, &j->work);
  napi_queue_async_work(env, j->work);
  return promise;
}
napi_value VerifyBatch(napi_env env, napi_callback_info info) {
  size_t argc = 5; napi_value a[5];
  napi_get_cb_info(env, info, &argc, a, nullptr, nullptr);
  Bytes sig, msgs, pks, dst; std::vector<uint64_t> off;
  if (argc < 5 || !get_bytes(env, a[0], &sig) || !get_bytes(env, a[1], &msgs) || !get_offsets(env, a[2], msgs.n, &off) ||
      !get_bytes(env, a[3], &pks) || !get_bytes(env, a[4], &dst))
    return nullptr;
  const size_t n = off.size() - 1;
  if (n == 0) return range(env, "Expected non-empty messages array");
  if (sig.n != 96) return range(env, "verifyBatch: the signature must be 96 bytes (compressed)");
  if (pks.n != n * 48) return range(env, "verifyBatch: public keys must be n x 48 bytes (compressed)");
  auto* j = new Job;
  j->a.assign(sig.p, sig.p + sig.n); j->msgs.assign(msgs.p, msgs.p + msgs.n); j->pks.assign(pks.p, pks.p + pks.n);
  j->dst.assign(dst.p, dst.p + dst.n); j->off = off;
  return queue(env, j, "bls381_verify_batch");
}
napi_value SignBatch(napi_env env, napi_callback_info info) {
  size_t argc = 4; napi_value a[4];
  napi_get_cb_info(env, info, &argc, a, nullptr, nullptr);
  Bytes sks, msgs, dst; std::vector<uint64_t> off;
  if (argc < 4 || !get_bytes(env, a[0], &sks) || !get_bytes(env, a[1], &msgs) || !get_offsets(env, a[2], msgs.n, &off) || !get_bytes(env, a[3], &dst))
    return nullptr;
  const size_t n = off.size() - 1;
  if (n == 0 || sks.n != n * 32) return range(env, "signBatch: private keys must be n x 32 bytes");
  auto* j = new Job;
  j->is_sign = true;
  j->a.assign(sks.p, sks.p + sks.n); j->msgs.assign(msgs.p, msgs.p + msgs.n); j->dst.assign(dst.p, dst.p + dst.n); j->off = off;
  return queue(env, j, "bls381_sign_batch");
}

// deviceCount() -> number of GPUs the addon shards verifyBatch over
napi_value DeviceCount(napi_env env, napi_callback_info) {
  napi_value v;
  napi_create_int32(env, bls381_device_count(), &v);
  return v;
}

#define EXPORT(name, fn) { napi_value f; napi_create_function(env, name, NAPI_AUTO_LENGTH, fn, nullptr, &f); napi_set_named_property(env, exports, name, f); }

}  // namespace

NAPI_MODULE_INIT() {
  uint32_t mask = 1;
  if (const char* e = getenv("BLS381_B200_DEVICES")) mask = static_cast<uint32_t>(strtoul(e, nullptr, 0));
  if (bls381_init_devices(mask ? mask : 1u, nullptr) != 0) { napi_throw_error(env, nullptr, bls381_last_error()); return exports; }
  EXPORT("pairingBatch", PairingBatch)
  EXPORT("millerProduct", MillerProduct)
  EXPORT("finalExpBatch", FinalExpBatch)
  EXPORT("fp12Product", Fp12Product)
  EXPORT("g1Decompress", (Decompress<48, 96, bls381_g1_decompress_batch>))
  EXPORT("g2Decompress", (Decompress<96, 192, bls381_g2_decompress_batch>))
  EXPORT("g1Validate", (Validate<96, bls381_g1_validate_batch>))
  EXPORT("g2Validate", (Validate<192, bls381_g2_validate_batch>))
  EXPORT("aggregateG1", (Aggregate<48, bls381_aggregate_g1>))
  EXPORT("aggregateG2", (Aggregate<96, bls381_aggregate_g2>))
  EXPORT("g1ScalarMul", (ScalarMul<96, bls381_g1_scalar_mul_batch>))
  EXPORT("g2ScalarMul", (ScalarMul<192, bls381_g2_scalar_mul_batch>))
  EXPORT("getPublicKeyBatch", GetPublicKeyBatch)
  EXPORT("hashToG2", (HashToCurve<192, bls381_hash_to_g2_batch>))
  EXPORT("hashToG1", (HashToCurve<96, bls381_hash_to_g1_batch>))
  EXPORT("g1FromUncompressed", (Decompress<96, 96, bls381_g1_from_uncompressed_batch>))
  EXPORT("g2FromUncompressed", (Decompress<192, 192, bls381_g2_from_uncompressed_batch>))
  EXPORT("g1Encode", (Encode<96, bls381_g1_encode_batch>))
  EXPORT("g2Encode", (Encode<192, bls381_g2_encode_batch>))
  EXPORT("verifyBatch", VerifyBatch)
  EXPORT("signBatch", SignBatch)
  EXPORT("deviceCount", DeviceCount)
  return exports;
}

// N-API shim over the C ABI (include/bls381_b200.h).  SOURCE ONLY: this image has no node / node_api.h, so the
// file is neither compiled nor tested here; the ctypes binding noble_bls12_381_b200/_lib.py is the
// binding exercised by the test-suite.  Build on a machine with node >= 18:
//   node-gyp configure build   (binding.gyp: sources ["addon.cc"], include_dirs ["../include"],
//                               libraries ["-L../noble_bls12_381_b200", "-lbls381_b200"])
// Synchronous entry points call straight through; the Promise-returning functions of the reference
// (sign / verify / verifyBatch, index.ts:746, 756, 792) are wrapped with napi_create_async_work so that the
// libuv thread blocks on the CUDA stream, never the JS thread.
#include <node_api.h>

#include <string>
#include <vector>

#include "bls381_b200.h"

namespace {

struct Bytes { uint8_t* p; size_t n; };

bool get_bytes(napi_env env, napi_value v, Bytes* out) {
  bool is_ta = false;
  napi_is_typedarray(env, v, &is_ta);
  if (!is_ta) { napi_throw_type_error(env, nullptr, "expected Uint8Array"); return false; }
  napi_typedarray_type t; napi_value ab; size_t off;
  napi_get_typedarray_info(env, v, &t, &out->n, reinterpret_cast<void**>(&out->p), &ab, &off);
  if (t != napi_uint8_array) { napi_throw_type_error(env, nullptr, "expected Uint8Array"); return false; }
  return true;
}

napi_value make_u8(napi_env env, size_t n, uint8_t** data) {
  napi_value ab, ta;
  napi_create_arraybuffer(env, n, reinterpret_cast<void**>(data), &ab);
  napi_create_typedarray(env, napi_uint8_array, n, ab, 0, &ta);
  return ta;
}

napi_value fail(napi_env env) { napi_throw_error(env, nullptr, bls381_last_error()); return nullptr; }

// pairingBatch(g1: Uint8Array(n*96), g2: Uint8Array(n*192), withFinalExponent: boolean) -> Uint8Array(n*576)
napi_value PairingBatch(napi_env env, napi_callback_info info) {
  size_t argc = 3; napi_value a[3];
  napi_get_cb_info(env, info, &argc, a, nullptr, nullptr);
  Bytes g1, g2; bool fe = true;
  if (!get_bytes(env, a[0], &g1) || !get_bytes(env, a[1], &g2)) return nullptr;
  napi_get_value_bool(env, a[2], &fe);
  const size_t n = g1.n / 96;
  if (g1.n != n * 96 || g2.n != n * 192) { napi_throw_range_error(env, nullptr, "bad buffer sizes"); return nullptr; }
  uint8_t* out; napi_value r = make_u8(env, n * 576, &out);
  if (bls381_pairing_batch(g1.p, g2.p, n, fe, out, nullptr) != 0) return fail(env);
  return r;
}

// millerProduct(g1, g2, withFinalExponent) -> Uint8Array(576)     (core of verify / verifyBatch)
napi_value MillerProduct(napi_env env, napi_callback_info info) {
  size_t argc = 3; napi_value a[3];
  napi_get_cb_info(env, info, &argc, a, nullptr, nullptr);
  Bytes g1, g2; bool fe = true;
  if (!get_bytes(env, a[0], &g1) || !get_bytes(env, a[1], &g2)) return nullptr;
  napi_get_value_bool(env, a[2], &fe);
  const size_t n = g1.n / 96;
  uint8_t* out; napi_value r = make_u8(env, 576, &out);
  if (bls381_miller_product(g1.p, g2.p, n, fe, out) != 0) return fail(env);
  return r;
}

// g1Decompress(keys: Uint8Array(n*48)) -> { points: Uint8Array(n*96), status: Int32Array(n) }   (PointG1.fromHex)
// g2Decompress(sigs: Uint8Array(n*96)) -> { points: Uint8Array(n*192), status: Int32Array(n) }  (PointG2.fromSignature)
template <int IN, int OUT, int (*FN)(const uint8_t*, size_t, uint8_t*, int32_t*)>
napi_value Decompress(napi_env env, napi_callback_info info) {
  size_t argc = 1; napi_value a[1];
  napi_get_cb_info(env, info, &argc, a, nullptr, nullptr);
  Bytes in;
  if (!get_bytes(env, a[0], &in)) return nullptr;
  const size_t n = in.n / IN;
  uint8_t* out; napi_value pts = make_u8(env, n * OUT, &out);
  napi_value ab, st; int32_t* sp;
  napi_create_arraybuffer(env, n * 4, reinterpret_cast<void**>(&sp), &ab);
  napi_create_typedarray(env, napi_int32_array, n, ab, 0, &st);
  if (FN(in.p, n, out, sp) != 0) return fail(env);
  napi_value obj; napi_create_object(env, &obj);
  napi_set_named_property(env, obj, "points", pts);
  napi_set_named_property(env, obj, "status", st);
  return obj;
}

// ---- async work: verifyBatch(sig96, packedMsgs, offsets(BigUint64Array n+1), pks(n*48), dst) -> Promise<{verdict, status}>
struct VerifyJob {
  napi_async_work work; napi_deferred deferred;
  std::vector<uint8_t> sig, msgs, pks, dst; std::vector<uint64_t> off; std::vector<int32_t> status;
  int verdict = 0, rc = 0; std::string err;
};
void VerifyExec(napi_env, void* d) {
  auto* j = static_cast<VerifyJob*>(d);
  const size_t n = j->off.size() - 1;
  j->status.resize(n + 1);
  j->rc = bls381_verify_batch(j->sig.data(), j->msgs.data(), j->off.data(), j->pks.data(), n, j->dst.data(), j->dst.size(),
                              &j->verdict, j->status.data());
  if (j->rc) j->err = bls381_last_error();
}
void VerifyDone(napi_env env, napi_status, void* d) {
  auto* j = static_cast<VerifyJob*>(d);
  if (j->rc) {
    napi_value msg, e; napi_create_string_utf8(env, j->err.c_str(), NAPI_AUTO_LENGTH, &msg); napi_create_error(env, nullptr, msg, &e);
    napi_reject_deferred(env, j->deferred, e);
  } else {
    napi_value obj, v, ab, st; int32_t* sp;
    napi_create_object(env, &obj);
    napi_create_int32(env, j->verdict, &v);
    napi_create_arraybuffer(env, j->status.size() * 4, reinterpret_cast<void**>(&sp), &ab);
    for (size_t i = 0; i < j->status.size(); ++i) sp[i] = j->status[i];
    napi_create_typedarray(env, napi_int32_array, j->status.size(), ab, 0, &st);
    napi_set_named_property(env, obj, "verdict", v);
    napi_set_named_property(env, obj, "status", st);
    napi_resolve_deferred(env, j->deferred, obj);
  }
  napi_delete_async_work(env, j->work);
  delete j;
}
napi_value VerifyBatch(napi_env env, napi_callback_info info) {
  size_t argc = 5; napi_value a[5];
  napi_get_cb_info(env, info, &argc, a, nullptr, nullptr);
  Bytes sig, msgs, pks, dst;
  if (!get_bytes(env, a[0], &sig) || !get_bytes(env, a[1], &msgs) || !get_bytes(env, a[3], &pks) || !get_bytes(env, a[4], &dst)) return nullptr;
  napi_typedarray_type t; size_t on; uint64_t* op; napi_value ab; size_t boff;
  napi_get_typedarray_info(env, a[2], &t, &on, reinterpret_cast<void**>(&op), &ab, &boff);
  auto* j = new VerifyJob;  // inputs are COPIED: the reference never retains caller buffers (index.ts:159-163)
  j->sig.assign(sig.p, sig.p + sig.n); j->msgs.assign(msgs.p, msgs.p + msgs.n); j->pks.assign(pks.p, pks.p + pks.n);
  j->dst.assign(dst.p, dst.p + dst.n); j->off.assign(op, op + on);
  napi_value promise, name;
  napi_create_promise(env, &j->deferred, &promise);
  napi_create_string_utf8(env, "bls381_verify_batch", NAPI_AUTO_LENGTH, &name);
  napi_create_async_work(env, nullptr, name, VerifyExec, VerifyDone, j, &j->work);
  napi_queue_async_work(env, j->work);
  return promise;
}

#define EXPORT(name, fn) { napi_value f; napi_create_function(env, name, NAPI_AUTO_LENGTH, fn, nullptr, &f); napi_set_named_property(env, exports, name, f); }

}  // namespace

NAPI_MODULE_INIT() {
  if (bls381_init(0, nullptr) != 0) { napi_throw_error(env, nullptr, bls381_last_error()); return exports; }
  EXPORT("pairingBatch", PairingBatch)
  EXPORT("millerProduct", MillerProduct)
  EXPORT("g1Decompress", (Decompress<48, 96, bls381_g1_decompress_batch>))
  EXPORT("g2Decompress", (Decompress<96, 192, bls381_g2_decompress_batch>))
  EXPORT("verifyBatch", VerifyBatch)
  // signBatch / aggregateG1 / aggregateG2 / hashToG2 / fp12Product / finalExpBatch follow the same two patterns.
  return exports;
}

/* Minimal DECLARATION STUB of Node's N-API (node_api.h / js_native_api.h) -- test infrastructure only.
 *
 * The build image has neither node nor its headers, so js/addon.cc cannot be compiled for real here.  This stub declares
 * exactly the types, enumerators and functions addon.cc uses, with the signatures Node documents for N-API version 8, so
 * that `g++ -fsyntax-only -I js/stub -I include js/addon.cc` (tests/test_js_shim.py) catches syntax and type errors in the
 * shim.  It is never used for a real build: node-gyp puts Node's own node_api.h first on the include path.
 */
#ifndef BLS381_STUB_NODE_API_H
#define BLS381_STUB_NODE_API_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct napi_env__* napi_env;
typedef struct napi_value__* napi_value;
typedef struct napi_callback_info__* napi_callback_info;
typedef struct napi_deferred__* napi_deferred;
typedef struct napi_async_work__* napi_async_work;

typedef enum { napi_ok = 0, napi_invalid_arg, napi_object_expected, napi_string_expected, napi_generic_failure = 9, napi_pending_exception = 10, napi_cancelled = 11 } napi_status;
typedef enum {
  napi_int8_array, napi_uint8_array, napi_uint8_clamped_array, napi_int16_array, napi_uint16_array, napi_int32_array,
  napi_uint32_array, napi_float32_array, napi_float64_array, napi_bigint64_array, napi_biguint64_array
} napi_typedarray_type;

typedef napi_value (*napi_callback)(napi_env env, napi_callback_info info);
typedef void (*napi_async_execute_callback)(napi_env env, void* data);
typedef void (*napi_async_complete_callback)(napi_env env, napi_status status, void* data);

#define NAPI_AUTO_LENGTH SIZE_MAX

napi_status napi_get_cb_info(napi_env env, napi_callback_info cbinfo, size_t* argc, napi_value* argv, napi_value* this_arg, void** data);
napi_status napi_is_typedarray(napi_env env, napi_value value, bool* result);
napi_status napi_get_typedarray_info(napi_env env, napi_value typedarray, napi_typedarray_type* type, size_t* length, void** data,
                                     napi_value* arraybuffer, size_t* byte_offset);
napi_status napi_create_arraybuffer(napi_env env, size_t byte_length, void** data, napi_value* result);
napi_status napi_create_typedarray(napi_env env, napi_typedarray_type type, size_t length, napi_value arraybuffer, size_t byte_offset,
                                   napi_value* result);
napi_status napi_get_value_bool(napi_env env, napi_value value, bool* result);
napi_status napi_create_int32(napi_env env, int32_t value, napi_value* result);
napi_status napi_create_object(napi_env env, napi_value* result);
napi_status napi_set_named_property(napi_env env, napi_value object, const char* utf8name, napi_value value);
napi_status napi_create_string_utf8(napi_env env, const char* str, size_t length, napi_value* result);
napi_status napi_create_error(napi_env env, napi_value code, napi_value msg, napi_value* result);
napi_status napi_throw_error(napi_env env, const char* code, const char* msg);
napi_status napi_throw_type_error(napi_env env, const char* code, const char* msg);
napi_status napi_throw_range_error(napi_env env, const char* code, const char* msg);
napi_status napi_create_function(napi_env env, const char* utf8name, size_t length, napi_callback cb, void* data, napi_value* result);
napi_status napi_create_promise(napi_env env, napi_deferred* deferred, napi_value* promise);
napi_status napi_resolve_deferred(napi_env env, napi_deferred deferred, napi_value resolution);
napi_status napi_reject_deferred(napi_env env, napi_deferred deferred, napi_value rejection);
napi_status napi_create_async_work(napi_env env, napi_value async_resource, napi_value async_resource_name,
                                   napi_async_execute_callback execute, napi_async_complete_callback complete, void* data,
                                   napi_async_work* result);
napi_status napi_queue_async_work(napi_env env, napi_async_work work);
napi_status napi_delete_async_work(napi_env env, napi_async_work work);

#ifdef __cplusplus
}
#define NAPI_MODULE_INIT() extern "C" napi_value napi_register_module_v1(napi_env env, napi_value exports)
#endif
#endif

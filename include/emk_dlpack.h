/* DLPack v0.x structure layout (the public in-memory tensor exchange ABI, dmlc/dlpack).
 * Only the structs libemk needs to READ a tensor that a framework exported with
 * `__dlpack__()` / `to_dlpack()`; libemk never allocates or frees these. */
#ifndef EMK_DLPACK_H_
#define EMK_DLPACK_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  kDLCPU = 1,
  kDLCUDA = 2,
  kDLCUDAHost = 3,
  kDLCUDAManaged = 13
} DLDeviceType;

typedef struct {
  int32_t device_type; /* DLDeviceType */
  int32_t device_id;
} DLDevice;

typedef enum { kDLInt = 0U, kDLUInt = 1U, kDLFloat = 2U, kDLBfloat = 4U } DLDataTypeCode;

typedef struct {
  uint8_t code;
  uint8_t bits;
  uint16_t lanes;
} DLDataType;

typedef struct {
  void* data;
  DLDevice device;
  int32_t ndim;
  DLDataType dtype;
  int64_t* shape;
  int64_t* strides; /* in elements; NULL => compact row-major */
  uint64_t byte_offset;
} DLTensor;

typedef struct DLManagedTensor {
  DLTensor dl_tensor;
  void* manager_ctx;
  void (*deleter)(struct DLManagedTensor* self);
} DLManagedTensor;

#ifdef __cplusplus
}
#endif
#endif /* EMK_DLPACK_H_ */

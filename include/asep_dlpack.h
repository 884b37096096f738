/*
 * asep_dlpack.h -- the subset of the DLPack v0.8 ABI that libasep.so exchanges.
 * Struct layouts follow the public DLPack specification (dmlc/dlpack, dlpack.h) so a
 * `DLTensor*` obtained from any DLPack producer (e.g. the DLManagedTensor inside the
 * capsule returned by torch.utils.dlpack.to_dlpack) can be passed unchanged.
 */
#ifndef ASEP_DLPACK_H_
#define ASEP_DLPACK_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef DLPACK_VERSION /* do not clash with a real dlpack.h included earlier */
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

typedef enum {
  kDLInt = 0U,
  kDLUInt = 1U,
  kDLFloat = 2U,
  kDLBfloat = 4U
} DLDataTypeCode;

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
  int64_t* strides; /* NULL = compact row-major */
  uint64_t byte_offset;
} DLTensor;

typedef struct DLManagedTensor {
  DLTensor dl_tensor;
  void* manager_ctx;
  void (*deleter)(struct DLManagedTensor* self);
} DLManagedTensor;
#endif

#ifdef __cplusplus
}
#endif
#endif /* ASEP_DLPACK_H_ */

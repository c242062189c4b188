// Shim (test infrastructure): the only macro uniq_hashtable.h needs from TensorFlow.
#pragma once
#define TF_DISALLOW_COPY_AND_ASSIGN(TypeName) \
  TypeName(const TypeName&) = delete;         \
  void operator=(const TypeName&) = delete

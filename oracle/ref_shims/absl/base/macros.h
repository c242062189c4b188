// Shim (test infrastructure).
#pragma once
#define ABSL_PREDICT_TRUE(x) (__builtin_expect(false || (x), true))
#define ABSL_PREDICT_FALSE(x) (__builtin_expect(false || (x), false))
#define GOOGLE_PREDICT_TRUE(x) ABSL_PREDICT_TRUE(x)
#define GOOGLE_PREDICT_FALSE(x) ABSL_PREDICT_FALSE(x)
#define ABSL_FALLTHROUGH_INTENDED [[fallthrough]]

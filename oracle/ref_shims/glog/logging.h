// Shim (test infrastructure): checks compiled out / minimal.
#pragma once
#include <cstdlib>
#include <iostream>
struct OrcNullStream { template <class T> OrcNullStream& operator<<(const T&) { return *this; } };
#define DCHECK(c) while (false) OrcNullStream()
#define DCHECK_LT(a, b) while (false) OrcNullStream()
#define DCHECK_LE(a, b) while (false) OrcNullStream()
#define DCHECK_GT(a, b) while (false) OrcNullStream()
#define DCHECK_GE(a, b) while (false) OrcNullStream()
#define DCHECK_EQ(a, b) while (false) OrcNullStream()
#define DCHECK_NE(a, b) while (false) OrcNullStream()
#define CHECK(c) if (!(c)) std::abort(); else OrcNullStream()
#define CHECK_LT(a, b) CHECK((a) < (b))
#define CHECK_LE(a, b) CHECK((a) <= (b))
#define CHECK_GT(a, b) CHECK((a) > (b))
#define CHECK_GE(a, b) CHECK((a) >= (b))
#define CHECK_EQ(a, b) CHECK((a) == (b))
#define CHECK_NE(a, b) CHECK((a) != (b))
#define LOG(x) OrcNullStream()

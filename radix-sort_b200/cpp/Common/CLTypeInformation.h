// CLTypeInformation.h -- key-type table (reference Common/CLTypeInformation.h:8-46).  The OpenCL spelling
// is kept for the CSV "Datatype" column's sibling uses; `key_type` is what the C ABI is told instead of a
// generated `#define DataType ...` preamble (reference src/RadixSortGPU.cpp:431-442).
#pragma once
#include "rsb200.h"

#include <cstdint>
#include <string_view>

template <typename T>
struct TypeNameString {};

#define RS_TYPE_ROW(T, CL, HAS_KEY, KEY)                                  \
    template <>                                                           \
    struct TypeNameString<T> {                                            \
        inline static constexpr std::string_view stdint_name  = #T;      \
        inline static constexpr std::string_view open_cl_name = CL;      \
        inline static constexpr bool             sortable     = HAS_KEY; \
        inline static constexpr rs_key_type      key_type     = KEY;     \
    };
RS_TYPE_ROW(int16_t, "short", true, RS_KEY_I16)            // named by the reference's table, not instantiated by its sorter
RS_TYPE_ROW(uint16_t, "unsigned short", true, RS_KEY_U16)  // (src/RadixSortGPU.cpp:598-601); here: two 8-bit passes
RS_TYPE_ROW(int32_t, "int", true, RS_KEY_I32)
RS_TYPE_ROW(uint32_t, "unsigned int", true, RS_KEY_U32)
RS_TYPE_ROW(int64_t, "long", true, RS_KEY_I64)
RS_TYPE_ROW(uint64_t, "unsigned long", true, RS_KEY_U64)
// additive (the reference excludes floating point, doc/doc.tex:64): ordered by value, NaNs last (rsb200.h rs_key_type)
RS_TYPE_ROW(float, "float", true, RS_KEY_F32)
RS_TYPE_ROW(double, "double", true, RS_KEY_F64)
#undef RS_TYPE_ROW

#!/usr/bin/perl -p
# TEST INFRASTRUCTURE: mechanical MSVC -> GCC token rewrites applied to a scratch copy of reference sources
# (never to the repo, never to /root/reference). No arithmetic is changed.
BEGIN { %T = (f64=>'double', f32=>'float', u64=>'uint64_t', i64=>'int64_t', u32=>'uint32_t', i32=>'int32_t',
              u16=>'uint16_t', i16=>'int16_t', u8=>'uint8_t', i8=>'int8_t'); }
# vector lane members:  x.m256d_f64[i]  ->  sr_lanes<double>(x)[i]   (also bare ".m256i_u64 + k" pointer uses)
s/((?:[A-Za-z_]\w*(?:\[[^\]]*\])?(?:\.|->))*[A-Za-z_]\w*(?:\[[^\]]*\])?)\.m(?:128|256)[di]?_(f64|f32|u64|i64|u32|i32|u16|i16|u8|i8)\b/sr_lanes<$T{$2}>($1)/g;
# integer literal suffixes
s/\b(0x[0-9a-fA-F]+|\d+)ui64\b/$1ULL/g;
s/\b(0x[0-9a-fA-F]+|\d+)i64\b/$1LL/g;
s/\b(0x[0-9a-fA-F]+|\d+)ui32\b/uint32_t($1)/g;
s/\b(0x[0-9a-fA-F]+|\d+)ui16\b/uint16_t($1)/g;
s/\b(0x[0-9a-fA-F]+|\d+)ui8\b/uint8_t($1)/g;
s/(?<![\w.])-(\d+)i8\b/int8_t(-$1)/g;
s/(?<![\w.])-(\d+)i32\b/int32_t(-$1)/g;
s/\b(\d+)i32\b/int32_t($1)/g;
s/\b(\d+)i8\b/int8_t($1)/g;
# reference typo inside a never-instantiated template (SRPacked64.h:40) that MSVC does not parse but g++ does
s/^(\s*)Packed64 ans;/$1SRPacked64 ans;/;
# MSVC accepts a member specialisation without 'template<>' (CETrainOperation.cpp:15)
s/^void CETrainOperation<SRDoubleNumber>::ProcessOne/template<> void CETrainOperation<SRDoubleNumber>::ProcessOne/;

// tor_oracle_video.hpp — CPU oracle for the video-export side of the render path:
// a restatement of trace_of_radiance/io/rgb.nim, io/color_conversions.nim and io/h264.nim.
//
// *** TEST INFRASTRUCTURE ONLY *** (same rules as tor_oracle.hpp).  Integer / byte work: the bar is bit-exact.
//
// Pins: the reference's own self-test for the colour conversion compares against yuv_rgb.c, which is NOT in the
// reference repository (color_conversions.nim:329-331), so there is no golden vector for it; the restatement is
// pinned by the BT.601 known answers the algorithm's source documents (white 235/128/128, black 16/128/128,
// red 81/90/240, ... — tests/test_video_export.py) and, for the H.264 stream, by decoding it with an independent
// decoder (FFmpeg through cv2) and by muxing it with the reference's own minimp4.h (oracle/_ref).
// Paths below are relative to /root/reference/trace_of_radiance/.
#pragma once
#include <stdint.h>
#include <string.h>

#include <vector>

namespace oracle {

// ---------------------------------------------------------------------------------- io/rgb.nim
// :20-21 `uint8(256 * clamp(c, 0.0, 0.999))`; NaN (UB in the reference) is defined as 0, like ppm_conv.
static inline uint8_t rgb_conv(double c) {
  if (c != c) return 0;
  double cl = c < 0.0 ? 0.0 : (c > 0.999 ? 0.999 : c);
  return (uint8_t)(int)(256 * cl);
}

// :17-31 `toRGB_Raw`.  As written the loop stores canvas row (nrows - i) into output row i (:29-31), which is one
// row off: output row 0 reads canvas row `nrows` (past the end of the buffer — undefined) and canvas row 0 never
// appears.  as_written = true reproduces that indexing with the out-of-bounds row defined as zeros;
// as_written = false is the evident intent, output row i = canvas row nrows-1-i (the PPM order of io/ppm.nim:20).
static inline void to_rgb_raw(const double* pixels, int32_t nrows, int32_t ncols, bool as_written, uint8_t* out) {
  for (int32_t i = nrows - 1; i >= 0; --i)
    for (int32_t j = 0; j < ncols; ++j) {
      const int32_t src_row = as_written ? nrows - i : nrows - 1 - i;
      uint8_t* o = out + 3 * ((int64_t)i * ncols + j);
      if (src_row >= nrows) {
        o[0] = o[1] = o[2] = 0;
        continue;
      }
      const double* p = pixels + 3 * ((int64_t)src_row * ncols + j);
      o[0] = rgb_conv(p[0]);
      o[1] = rgb_conv(p[1]);
      o[2] = rgb_conv(p[2]);
    }
}

// ---------------------------------------------------------------------- io/color_conversions.nim
struct YCbCrCoefs {  // :66-73
  uint8_t kr, kg, kb, fb, fr, y_scale, y_min;
};
static inline uint8_t to_fixed_u8(double x, int precision) {  // :101-102
  return (uint8_t)(x * (double)(1 << precision) + 0.5);
}
static inline YCbCrCoefs bt601_coefs() {  // :104-114 with the arguments of :178 (0.299, 0.114, 16, 235, 240-16)
  const double kr = 0.299, kb = 0.114, ymin = 16.0, ymax = 235.0, range = 240.0 - 16.0;
  YCbCrCoefs c;
  c.kr = to_fixed_u8(kr, 8);
  c.kb = to_fixed_u8(kb, 8);
  c.kg = (uint8_t)(256 - c.kr - c.kb);
  c.fb = to_fixed_u8((range / 255.0) / (2.0 * (1.0 - kb)), 8);
  c.fr = to_fixed_u8((range / 255.0) / (2.0 * (1.0 - kr)), 8);
  c.y_scale = to_fixed_u8((ymax - ymin) / 255.0, 7);
  c.y_min = (uint8_t)ymin;
  return c;
}

// :180-252 `rgbRaw_to_ycbcr420`: packed RGB (RGB RGB ...) -> planar Y' (width x height), Cb, Cr (half size each,
// stride (width+1)/2, :131).  Width and height must be even (:201-202).  uint16 / int16 intermediates as in the
// reference; `shr` on a signed Nim integer is an arithmetic shift.
static inline bool rgb_to_ycbcr420(int32_t width, int32_t height, const uint8_t* rgb, uint8_t* Y, uint8_t* U,
                                   uint8_t* V) {
  if ((width & 1) || (height & 1)) return false;
  const YCbCrCoefs k = bt601_coefs();
  const int32_t cstride = (width + 1) / 2;
  for (int32_t ii = 0; ii < height; ii += 2)
    for (int32_t jj = 0; jj < width; jj += 2) {
      int16_t tU = 0, tV = 0;
      for (int32_t di = 0; di < 2; ++di)      // the reference unrolls (ii,jj) (ii,jj+1) (ii+1,jj) (ii+1,jj+1);
        for (int32_t dj = 0; dj < 2; ++dj) {  // the accumulation is integer, so the order is immaterial
          const uint8_t* p = rgb + 3 * ((int64_t)(ii + di) * width + (jj + dj));
          const uint16_t tY = (uint16_t)((uint16_t)((uint16_t)k.kr * p[0] + (uint16_t)k.kg * p[1] + (uint16_t)k.kb * p[2]) >> 8);
          tU = (int16_t)(tU + ((int16_t)p[2] - (int16_t)tY));
          tV = (int16_t)(tV + ((int16_t)p[0] - (int16_t)tY));
          Y[(int64_t)(ii + di) * width + (jj + dj)] = (uint8_t)((uint8_t)((uint16_t)(tY * k.y_scale) >> 7) + k.y_min);
        }
      const int16_t u = (int16_t)((int16_t)((int16_t)(tU >> 2) * (int16_t)k.fb) >> 8);
      const int16_t v = (int16_t)((int16_t)((int16_t)(tV >> 2) * (int16_t)k.fr) >> 8);
      U[(int64_t)(ii >> 1) * cstride + (jj >> 1)] = (uint8_t)(u + 128);
      V[(int64_t)(ii >> 1) * cstride + (jj >> 1)] = (uint8_t)(v + 128);
    }
  return true;
}

// ---------------------------------------------------------------------------------- io/h264.nim
// The "lossless" writer: a hand-built SPS, a constant PPS, and every macroblock as I_PCM (raw samples).
struct BitBuffer {  // :14-18
  int shift = 0;
  uint32_t cache = 0;
  uint8_t* buf = nullptr;
  int cursor = 0;
};
static inline void store_be32(uint8_t* dst, uint32_t v) {
  dst[0] = (uint8_t)(v >> 24);
  dst[1] = (uint8_t)(v >> 16);
  dst[2] = (uint8_t)(v >> 8);
  dst[3] = (uint8_t)v;
}
static inline void bb_put(BitBuffer& bb, int n, uint32_t val) {  // :58-69
  bb.shift -= n;
  if (bb.shift < 0) {
    bb.cache |= val >> -bb.shift;
    store_be32(bb.buf + bb.cursor, bb.cache);
    bb.cursor += 4;
    bb.shift += 32;
    bb.cache = 0;
  }
  bb.cache |= val << bb.shift;
}
static inline void bb_put_golomb(BitBuffer& bb, uint32_t val) {  // :71-76
  int size = 1;
  uint32_t t = val + 1;
  while ((t >>= 1) != 0) ++size;
  bb_put(bb, 2 * size - 1, val + 1);
}
static inline void bb_flush(BitBuffer& bb) {  // :78-80
  store_be32(bb.buf + bb.cursor, bb.cache);
  bb.cursor += 4;
}

// :88-142 `initSPS` (needCropping is never set by the reference, :156 "TODO cropping": always false)
static inline std::vector<uint8_t> h264_sps(int width, int height) {
  std::vector<uint8_t> sps(40, 0);
  BitBuffer bb;
  bb.buf = sps.data();
  sps[3] = 0x01;  // start code 00 00 00 01 (:100)
  bb.shift = 32;
  bb.cursor = 4;
  bb_put(bb, 1, 0);   // forbidden_zero_bit
  bb_put(bb, 2, 3);   // nal_ref_idc
  bb_put(bb, 5, 7);   // nal_unit_type = SPS
  bb_put(bb, 8, 66);  // Baseline profile
  bb_put(bb, 1, 0);   // constraint_set0..3
  bb_put(bb, 1, 0);
  bb_put(bb, 1, 0);
  bb_put(bb, 1, 0);
  bb_put(bb, 4, 0);   // reserved_zero_4bits
  bb_put(bb, 8, 10);  // level_idc
  bb_put_golomb(bb, 0);  // seq_parameter_set_id
  bb_put_golomb(bb, 0);  // log2_max_frame_num_minus4
  bb_put_golomb(bb, 0);  // pic_order_cnt_type
  bb_put_golomb(bb, 0);  // log2_max_pic_order_cnt_lsb_minus4
  bb_put_golomb(bb, 0);  // num_ref_frames
  bb_put(bb, 1, 0);      // gaps_in_frame_num_value_allowed_flag
  bb_put_golomb(bb, (uint32_t)(((width + 15) >> 4) - 1));   // pic_width_in_mbs_minus_1
  bb_put_golomb(bb, (uint32_t)(((height + 15) >> 4) - 1));  // pic_height_in_map_units_minus_1
  bb_put(bb, 1, 1);  // frame_mbs_only_flag
  bb_put(bb, 1, 0);  // direct_8x8_inference_flag
  bb_put(bb, 1, 0);  // frame_cropping_flag
  bb_put(bb, 1, 0);  // vui_parameters_present_flag
  bb_put(bb, 1, 1);  // stop bit
  bb_flush(bb);
  sps.resize((size_t)(bb.cursor - bb.shift / 8));  // :142
  return sps;
}

static const uint8_t kH264Pps[8] = {0x00, 0x00, 0x00, 0x01, 0x68, 0xce, 0x38, 0x80};                  // :36
static const uint8_t kH264SliceHeader[9] = {0x00, 0x00, 0x00, 0x01, 0x05, 0x88, 0x84, 0x21, 0xa0};  // :37
static const uint8_t kH264MacroblockHeader[2] = {0x0d, 0x00};                                          // :38
static const uint8_t kH264SliceStopBit = 0x80;                                                         // :39

// `H264Encoder.init` (:159-168): the stream header = SPS + PPS
static inline void h264_header(int width, int height, std::vector<uint8_t>& out) {
  std::vector<uint8_t> sps = h264_sps(width, height);
  out.insert(out.end(), sps.begin(), sps.end());
  out.insert(out.end(), kH264Pps, kH264Pps + 8);
}

// `flushFrame` (:249-259) + `encodeMacroblock` (:189-202): slice header, then for every 16x16 macroblock in raster
// order [header except for the first] 256 luma, 64 Cb, 64 Cr raw samples, then the stop byte.  Only whole macroblocks
// are written (lumaHeight div 16, lumaWidth div 16); the chroma row stride is lumaWidth shr 1 (:48-52).
static inline void h264_frame(int width, int height, const uint8_t* Y, const uint8_t* Cb, const uint8_t* Cr,
                              std::vector<uint8_t>& out) {
  out.insert(out.end(), kH264SliceHeader, kH264SliceHeader + 9);
  const int cw = width >> 1;
  for (int i = 0; i < height / 16; ++i)
    for (int j = 0; j < width / 16; ++j) {
      if (!(i == 0 && j == 0)) out.insert(out.end(), kH264MacroblockHeader, kH264MacroblockHeader + 2);
      for (int x = i * 16; x < (i + 1) * 16; ++x)
        for (int y = j * 16; y < (j + 1) * 16; ++y) out.push_back(Y[(int64_t)x * width + y]);
      for (int x = i * 8; x < (i + 1) * 8; ++x)
        for (int y = j * 8; y < (j + 1) * 8; ++y) out.push_back(Cb[(int64_t)x * cw + y]);
      for (int x = i * 8; x < (i + 1) * 8; ++x)
        for (int y = j * 8; y < (j + 1) * 8; ++y) out.push_back(Cr[(int64_t)x * cw + y]);
    }
  out.push_back(kH264SliceStopBit);
}

}  // namespace oracle

// tor_video.cc — host side of the animation export: the I_PCM H.264 elementary-stream
// writer that io/h264.nim implements and the MP4 muxing step of io/mp4.nim.  Pure byte formatting and file I/O; the
// arithmetic that feeds it (render, draw, RGB quantisation, BT.601 4:2:0 conversion) runs on the device
// (tor_render_ycbcr420, tor_api.cu).
//
// Every header byte is produced from its H.264 syntax elements with a bit writer, so the stream can be read against
// ITU-T H.264 §7.3; tests/test_video_export.py checks that the result is byte-identical to the oracle's restatement of
// io/h264.nim (whose SPS / PPS / slice header are the reference's constants) and decodes it with FFmpeg.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <algorithm>
#include <vector>

#include "../../include/tor_b200.h"

namespace {

// ------------------------------------------------------------------------------ H.264 bit syntax
class BitWriter {  // MSB first, as H.264 §7.2 reads
 public:
  void u(int nbits, uint32_t value) {
    for (int b = nbits - 1; b >= 0; --b) bit((value >> b) & 1u);
  }
  void ue(uint32_t value) {  // Exp-Golomb, §9.1: (leading zeros) 1 (info bits)
    const uint32_t code = value + 1;
    int len = 0;
    while ((code >> len) > 1) ++len;
    u(len, 0);
    u(len + 1, code);
  }
  void se(int32_t value) { ue(value > 0 ? (uint32_t)(2 * value - 1) : (uint32_t)(-2 * value)); }
  void trailing_bits() {  // rbsp_trailing_bits: stop bit + alignment
    bit(1);
    align_zero();
  }
  void align_zero() {
    while (nbits_ & 7) bit(0);
  }
  const std::vector<uint8_t>& bytes() const { return bytes_; }

 private:
  void bit(uint32_t b) {
    if ((nbits_ & 7) == 0) bytes_.push_back(0);
    if (b) bytes_.back() |= (uint8_t)(0x80u >> (nbits_ & 7));
    ++nbits_;
  }
  std::vector<uint8_t> bytes_;
  size_t nbits_ = 0;
};

void start_code(std::vector<uint8_t>& out) {
  static const uint8_t sc[4] = {0, 0, 0, 1};
  out.insert(out.end(), sc, sc + 4);
}
void nal_header(BitWriter& w, uint32_t ref_idc, uint32_t type) {
  w.u(1, 0);  // forbidden_zero_bit
  w.u(2, ref_idc);
  w.u(5, type);
}

// Sequence parameter set, the field values of io/h264.nim:102-139: Baseline, level 1.0, frame numbers and picture
// order counts on 4 bits, no reference frames, frames only, no cropping, no VUI.
std::vector<uint8_t> make_sps(int width, int height) {
  BitWriter w;
  nal_header(w, 3, 7);
  w.u(8, 66);  // profile_idc: Baseline
  w.u(4, 0);   // constraint_set0..3_flag
  w.u(4, 0);   // reserved_zero_4bits
  w.u(8, 10);  // level_idc
  w.ue(0);     // seq_parameter_set_id
  w.ue(0);     // log2_max_frame_num_minus4
  w.ue(0);     // pic_order_cnt_type
  w.ue(0);     // log2_max_pic_order_cnt_lsb_minus4
  w.ue(0);     // max_num_ref_frames
  w.u(1, 0);   // gaps_in_frame_num_value_allowed_flag
  w.ue((uint32_t)((width + 15) / 16 - 1));   // pic_width_in_mbs_minus1
  w.ue((uint32_t)((height + 15) / 16 - 1));  // pic_height_in_map_units_minus1
  w.u(1, 1);   // frame_mbs_only_flag
  w.u(1, 0);   // direct_8x8_inference_flag
  // The reference leaves cropping as a TODO (io/h264.nim:168) and its stream for a size that is not a multiple of 16
  // lacks the last macroblock row / column.  Here the picture is coded in whole macroblocks (the edge samples
  // repeated) and cropped back: offsets are in units of two luma samples for 4:2:0 frames (H.264 7.4.2.1.1).
  const int crop_r = (((width + 15) / 16) * 16 - width) / 2, crop_b = (((height + 15) / 16) * 16 - height) / 2;
  if (crop_r || crop_b) {
    w.u(1, 1);  // frame_cropping_flag
    w.ue(0);    // frame_crop_left_offset
    w.ue((uint32_t)crop_r);
    w.ue(0);    // frame_crop_top_offset
    w.ue((uint32_t)crop_b);
  } else {
    w.u(1, 0);  // frame_cropping_flag (sizes that are multiples of 16: the reference's SPS byte for byte)
  }
  w.u(1, 0);   // vui_parameters_present_flag
  w.trailing_bits();
  return w.bytes();
}

// Picture parameter set (the constant of io/h264.nim:36): CAVLC, one slice group, QP 26, no deblocking control.
std::vector<uint8_t> make_pps() {
  BitWriter w;
  nal_header(w, 3, 8);
  w.ue(0);     // pic_parameter_set_id
  w.ue(0);     // seq_parameter_set_id
  w.u(1, 0);   // entropy_coding_mode_flag
  w.u(1, 0);   // bottom_field_pic_order_in_frame_present_flag
  w.ue(0);     // num_slice_groups_minus1
  w.ue(0);     // num_ref_idx_l0_default_active_minus1
  w.ue(0);     // num_ref_idx_l1_default_active_minus1
  w.u(1, 0);   // weighted_pred_flag
  w.u(2, 0);   // weighted_bipred_idc
  w.se(0);     // pic_init_qp_minus26
  w.se(0);     // pic_init_qs_minus26
  w.se(0);     // chroma_qp_index_offset
  w.u(1, 0);   // deblocking_filter_control_present_flag
  w.u(1, 0);   // constrained_intra_pred_flag
  w.u(1, 0);   // redundant_pic_cnt_present_flag
  w.trailing_bits();
  return w.bytes();
}

constexpr uint32_t kMbTypeIPCM = 25;  // Table 7-11

// IDR slice header (io/h264.nim:37) up to and including the first macroblock's mb_type and PCM alignment.
// nal_ref_idc is 0 as in the reference, so there is no dec_ref_pic_marking().
std::vector<uint8_t> make_slice_prefix() {
  BitWriter w;
  nal_header(w, 0, 5);
  w.ue(0);    // first_mb_in_slice
  w.ue(7);    // slice_type: I, all slices of the picture
  w.ue(0);    // pic_parameter_set_id
  w.u(4, 0);  // frame_num
  w.ue(0);    // idr_pic_id
  w.u(4, 0);  // pic_order_cnt_lsb
  w.se(0);    // slice_qp_delta
  w.ue(kMbTypeIPCM);
  w.align_zero();  // pcm_alignment_zero_bit
  return w.bytes();
}
std::vector<uint8_t> make_mb_prefix() {  // io/h264.nim:38
  BitWriter w;
  w.ue(kMbTypeIPCM);
  w.align_zero();
  return w.bytes();
}

// --------------------------------------------------------------------------------- ISO-BMFF boxes
struct Box {
  std::vector<uint8_t> b;
  void u8(uint32_t v) { b.push_back((uint8_t)v); }
  void u16(uint32_t v) {
    u8(v >> 8);
    u8(v);
  }
  void u32(uint32_t v) {
    u16(v >> 16);
    u16(v);
  }
  void u64(uint64_t v) {
    u32((uint32_t)(v >> 32));
    u32((uint32_t)v);
  }
  void tag(const char* t) { b.insert(b.end(), t, t + 4); }
  void zeros(int n) { b.insert(b.end(), (size_t)n, 0); }
  void raw(const std::vector<uint8_t>& v) { b.insert(b.end(), v.begin(), v.end()); }
  size_t open(const char* t) {  // returns the position of the size field
    size_t at = b.size();
    u32(0);
    tag(t);
    return at;
  }
  size_t open_full(const char* t, uint32_t version, uint32_t flags) {
    size_t at = open(t);
    u32((version << 24) | flags);
    return at;
  }
  void close(size_t at) {
    uint32_t n = (uint32_t)(b.size() - at);
    b[at] = (uint8_t)(n >> 24);
    b[at + 1] = (uint8_t)(n >> 16);
    b[at + 2] = (uint8_t)(n >> 8);
    b[at + 3] = (uint8_t)n;
  }
};

void unity_matrix(Box& x) {
  const uint32_t m[9] = {0x00010000, 0, 0, 0, 0x00010000, 0, 0, 0, 0x40000000};
  for (uint32_t v : m) x.u32(v);
}

struct NalSpan {
  size_t begin, end;  // payload without the start code
};

// Annex B byte stream -> NAL units (what get_nal_size / write_mp4 of io/mp4.nim:66-96 walk through).
std::vector<NalSpan> split_annexb(const std::vector<uint8_t>& d) {
  std::vector<NalSpan> out;
  const size_t n = d.size();
  size_t i = 0;
  auto start_at = [&](size_t p) -> int {  // length of a start code at p, or 0
    if (p + 3 <= n && d[p] == 0 && d[p + 1] == 0 && d[p + 2] == 1) return 3;
    if (p + 4 <= n && d[p] == 0 && d[p + 1] == 0 && d[p + 2] == 0 && d[p + 3] == 1) return 4;
    return 0;
  };
  while (i < n) {
    int sc = start_at(i);
    if (!sc) {
      ++i;
      continue;
    }
    size_t b = i + (size_t)sc, e = b;
    while (e < n && !start_at(e)) ++e;
    if (e > b) out.push_back(NalSpan{b, e});
    i = e;
  }
  return out;
}

}  // namespace

struct tor_h264_encoder {
  FILE* f = nullptr;
  int32_t width = 0, height = 0;
  uint8_t* frame = nullptr;  // Y (w*h) | Cb (w/2*h/2) | Cr, contiguous like Frame.buffer of io/h264.nim:20-26
  bool frame_pinned = false;
  int64_t frame_bytes = 0;
  std::vector<uint8_t> slice_prefix, mb_prefix, scratch;
};

extern "C" {

int tor_h264_open(const char* path, int32_t width, int32_t height, tor_h264_encoder** out) {
  if (!out) return TOR_ERR_INVALID_ARG;
  *out = nullptr;
  // 4:2:0 needs even sizes (color_conversions.nim:201-202); sizes that are not multiples of 16 are padded to whole
  // macroblocks and cropped in the SPS (make_sps) — the reference's own 576x324 preset
  // (trace_of_radiance_animation.nim:122-123) is such a size
  if (!path || width <= 0 || height <= 0 || (width % 2) || (height % 2)) return TOR_ERR_INVALID_ARG;
  FILE* f = fopen(path, "wb");
  if (!f) return TOR_ERR_INVALID_ARG;
  tor_h264_encoder* e = new tor_h264_encoder();
  e->f = f;
  e->width = width;
  e->height = height;
  e->frame_bytes = (int64_t)width * height + 2 * (int64_t)(width / 2) * (height / 2);
  e->frame = (uint8_t*)tor_host_alloc((size_t)e->frame_bytes);  // page-locked: the device writes frames straight into it
  e->frame_pinned = e->frame != nullptr;
  if (!e->frame) e->frame = (uint8_t*)malloc((size_t)e->frame_bytes);
  if (!e->frame) {
    fclose(f);
    delete e;
    return TOR_ERR_INVALID_ARG;
  }
  memset(e->frame, 0, (size_t)e->frame_bytes);
  e->slice_prefix = make_slice_prefix();
  e->mb_prefix = make_mb_prefix();
  std::vector<uint8_t> hdr;
  start_code(hdr);
  const std::vector<uint8_t> sps = make_sps(width, height), pps = make_pps();
  hdr.insert(hdr.end(), sps.begin(), sps.end());
  start_code(hdr);
  hdr.insert(hdr.end(), pps.begin(), pps.end());
  if (fwrite(hdr.data(), 1, hdr.size(), f) != hdr.size()) {
    tor_h264_finish(e);
    return TOR_ERR_INVALID_ARG;
  }
  *out = e;
  return TOR_OK;
}

int tor_h264_frame_buffer(tor_h264_encoder* e, uint8_t** buf, int64_t* size) {
  if (!e || !buf || !size) return TOR_ERR_INVALID_ARG;
  *buf = e->frame;
  *size = e->frame_bytes;
  return TOR_OK;
}

int tor_h264_flush_frame(tor_h264_encoder* e) {
  if (!e || !e->f) return TOR_ERR_INVALID_ARG;
  const int w = e->width, h = e->height, cw = w / 2;
  const uint8_t* Y = e->frame;
  const uint8_t* Cb = Y + (size_t)w * h;
  const uint8_t* Cr = Cb + (size_t)cw * (h / 2);
  std::vector<uint8_t>& s = e->scratch;
  s.clear();
  const int mbw = (w + 15) / 16, mbh = (h + 15) / 16, ch = h / 2;
  s.reserve((size_t)mbw * mbh * 386 + 16);
  start_code(s);
  s.insert(s.end(), e->slice_prefix.begin(), e->slice_prefix.end());
  for (int my = 0; my < mbh; ++my)
    for (int mx = 0; mx < mbw; ++mx) {
      if (my || mx) s.insert(s.end(), e->mb_prefix.begin(), e->mb_prefix.end());
      const bool inside = (mx + 1) * 16 <= w && (my + 1) * 16 <= h;
      if (inside) {
        for (int r = 0; r < 16; ++r) {  // pcm_sample_luma[256], raster order inside the macroblock
          const uint8_t* src = Y + (size_t)(my * 16 + r) * w + mx * 16;
          s.insert(s.end(), src, src + 16);
        }
        for (int r = 0; r < 8; ++r) {  // pcm_sample_chroma: Cb then Cr
          const uint8_t* src = Cb + (size_t)(my * 8 + r) * cw + mx * 8;
          s.insert(s.end(), src, src + 8);
        }
        for (int r = 0; r < 8; ++r) {
          const uint8_t* src = Cr + (size_t)(my * 8 + r) * cw + mx * 8;
          s.insert(s.end(), src, src + 8);
        }
      } else {  // a macroblock that sticks out of the picture: repeat the edge samples (cropped away by the SPS)
        for (int r = 0; r < 16; ++r)
          for (int c = 0; c < 16; ++c)
            s.push_back(Y[(size_t)std::min(my * 16 + r, h - 1) * w + std::min(mx * 16 + c, w - 1)]);
        for (const uint8_t* P : {Cb, Cr})
          for (int r = 0; r < 8; ++r)
            for (int c = 0; c < 8; ++c)
              s.push_back(P[(size_t)std::min(my * 8 + r, ch - 1) * cw + std::min(mx * 8 + c, cw - 1)]);
      }
    }
  s.push_back(0x80);  // rbsp_slice_trailing_bits: the data is byte aligned, so stop bit + 7 zeros
  return fwrite(s.data(), 1, s.size(), e->f) == s.size() ? TOR_OK : TOR_ERR_INVALID_ARG;
}

int tor_h264_finish(tor_h264_encoder* e) {
  if (!e) return TOR_ERR_INVALID_ARG;
  int rc = TOR_OK;
  if (e->f && fclose(e->f) != 0) rc = TOR_ERR_INVALID_ARG;
  if (e->frame) {
    if (e->frame_pinned)
      tor_host_free(e->frame);
    else
      free(e->frame);
  }
  delete e;
  return rc;
}

// MP4Muxer.initialize + writeMP4_from + close (io/mp4.nim:104-163): every parameter-set NAL goes to the avcC record,
// every other NAL becomes one length-prefixed sample lasting 90000 / fps ticks.  Layout: ftyp, mdat, moov.
int tor_mp4_mux_h264_file(const char* src_264, const char* dst_mp4, int32_t width, int32_t height, int32_t fps) {
  if (!src_264 || !dst_mp4 || width <= 0 || height <= 0 || fps <= 0 || fps > 90000) return TOR_ERR_INVALID_ARG;
  std::vector<uint8_t> es;
  {
    FILE* f = fopen(src_264, "rb");
    if (!f) return TOR_ERR_INVALID_ARG;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    es.resize(n > 0 ? (size_t)n : 0);
    size_t got = es.empty() ? 0 : fread(es.data(), 1, es.size(), f);
    fclose(f);
    if (got != es.size()) return TOR_ERR_INVALID_ARG;
  }
  const std::vector<NalSpan> nals = split_annexb(es);
  std::vector<uint8_t> sps, pps;
  std::vector<NalSpan> samples;
  for (const NalSpan& s : nals) {
    const uint32_t type = es[s.begin] & 0x1fu;
    if (type == 7) {
      if (sps.empty()) sps.assign(es.begin() + (long)s.begin, es.begin() + (long)s.end);
    } else if (type == 8) {
      if (pps.empty()) pps.assign(es.begin() + (long)s.begin, es.begin() + (long)s.end);
    } else {
      samples.push_back(s);
    }
  }
  if (sps.size() < 4 || pps.empty() || samples.empty()) return TOR_ERR_INVALID_ARG;

  FILE* out = fopen(dst_mp4, "wb");
  if (!out) return TOR_ERR_INVALID_ARG;
  bool ok = true;
  auto put = [&](const std::vector<uint8_t>& v) { ok = ok && fwrite(v.data(), 1, v.size(), out) == v.size(); };

  Box ftyp;
  size_t a = ftyp.open("ftyp");
  ftyp.tag("isom");
  ftyp.u32(0x200);
  ftyp.tag("isom");
  ftyp.tag("iso2");
  ftyp.tag("avc1");
  ftyp.tag("mp41");
  ftyp.close(a);
  put(ftyp.b);

  uint64_t payload = 0;
  for (const NalSpan& s : samples) payload += 4 + (s.end - s.begin);
  Box mdat;  // 64-bit box size so that long animations need no special case
  mdat.u32(1);
  mdat.tag("mdat");
  mdat.u64(16 + payload);
  put(mdat.b);
  uint64_t offset = ftyp.b.size() + mdat.b.size();
  std::vector<uint64_t> offsets;
  std::vector<uint32_t> sizes;
  for (const NalSpan& s : samples) {
    const uint32_t n = (uint32_t)(s.end - s.begin);
    const uint8_t len[4] = {(uint8_t)(n >> 24), (uint8_t)(n >> 16), (uint8_t)(n >> 8), (uint8_t)n};
    ok = ok && fwrite(len, 1, 4, out) == 4 && fwrite(es.data() + s.begin, 1, n, out) == n;
    offsets.push_back(offset);
    sizes.push_back(n + 4);
    offset += n + 4;
  }

  const uint32_t timescale = 90000, delta = timescale / (uint32_t)fps;  // io/mp4.nim:91 `90000 div fps`
  const uint64_t duration = (uint64_t)delta * samples.size();
  Box m;
  size_t moov = m.open("moov");
  {
    size_t b = m.open_full("mvhd", 0, 0);
    m.u32(0);  // creation, modification
    m.u32(0);
    m.u32(timescale);
    m.u32((uint32_t)duration);
    m.u32(0x00010000);  // rate 1.0
    m.u16(0x0100);      // volume
    m.zeros(10);
    unity_matrix(m);
    m.zeros(24);
    m.u32(2);  // next_track_ID
    m.close(b);
  }
  size_t trak = m.open("trak");
  {
    size_t b = m.open_full("tkhd", 0, 7);  // enabled, in movie, in preview
    m.u32(0);
    m.u32(0);
    m.u32(1);  // track_ID
    m.u32(0);
    m.u32((uint32_t)duration);
    m.zeros(8);
    m.u16(0);  // layer
    m.u16(0);  // alternate group
    m.u16(0);  // volume (video)
    m.u16(0);
    unity_matrix(m);
    m.u32((uint32_t)width << 16);
    m.u32((uint32_t)height << 16);
    m.close(b);
  }
  size_t mdia = m.open("mdia");
  {
    size_t b = m.open_full("mdhd", 0, 0);
    m.u32(0);
    m.u32(0);
    m.u32(timescale);
    m.u32((uint32_t)duration);
    m.u16(0x55c4);  // language: und
    m.u16(0);
    m.close(b);
    b = m.open_full("hdlr", 0, 0);
    m.u32(0);
    m.tag("vide");
    m.zeros(12);
    const char name[] = "VideoHandler";
    m.b.insert(m.b.end(), name, name + sizeof(name));
    m.close(b);
  }
  size_t minf = m.open("minf");
  {
    size_t b = m.open_full("vmhd", 0, 1);
    m.zeros(8);
    m.close(b);
    size_t dinf = m.open("dinf");
    size_t dref = m.open_full("dref", 0, 0);
    m.u32(1);
    size_t url = m.open_full("url ", 0, 1);  // media data in the same file
    m.close(url);
    m.close(dref);
    m.close(dinf);
  }
  size_t stbl = m.open("stbl");
  {
    size_t stsd = m.open_full("stsd", 0, 0);
    m.u32(1);
    size_t avc1 = m.open("avc1");
    m.zeros(6);
    m.u16(1);  // data_reference_index
    m.zeros(16);
    m.u16((uint32_t)width);
    m.u16((uint32_t)height);
    m.u32(0x00480000);  // 72 dpi
    m.u32(0x00480000);
    m.u32(0);
    m.u16(1);  // frame_count
    m.zeros(32);  // compressorname
    m.u16(0x0018);  // depth
    m.u16(0xffff);  // pre_defined = -1
    size_t avcc = m.open("avcC");
    m.u8(1);       // configurationVersion
    m.u8(sps[1]);  // AVCProfileIndication
    m.u8(sps[2]);  // profile_compatibility
    m.u8(sps[3]);  // AVCLevelIndication
    m.u8(0xff);    // 6 reserved bits + lengthSizeMinusOne = 3
    m.u8(0xe1);    // 3 reserved bits + one SPS
    m.u16((uint32_t)sps.size());
    m.raw(sps);
    m.u8(1);
    m.u16((uint32_t)pps.size());
    m.raw(pps);
    m.close(avcc);
    m.close(avc1);
    m.close(stsd);

    size_t b = m.open_full("stts", 0, 0);
    m.u32(1);
    m.u32((uint32_t)samples.size());
    m.u32(delta);
    m.close(b);
    b = m.open_full("stsc", 0, 0);  // one sample per chunk
    m.u32(1);
    m.u32(1);
    m.u32(1);
    m.u32(1);
    m.close(b);
    b = m.open_full("stsz", 0, 0);
    m.u32(0);
    m.u32((uint32_t)sizes.size());
    for (uint32_t v : sizes) m.u32(v);
    m.close(b);
    b = m.open_full("co64", 0, 0);
    m.u32((uint32_t)offsets.size());
    for (uint64_t v : offsets) m.u64(v);
    m.close(b);
    // no stss: every sample is an IDR picture, i.e. a sync sample
  }
  m.close(stbl);
  m.close(minf);
  m.close(mdia);
  m.close(trak);
  m.close(moov);
  put(m.b);
  if (fclose(out) != 0) ok = false;
  return ok ? TOR_OK : TOR_ERR_INVALID_ARG;
}

}  // extern "C"

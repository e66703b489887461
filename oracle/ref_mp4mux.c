/* ref_mp4mux.c — drives the REFERENCE's own MP4 muxer, the vendored C header
 * /root/reference/trace_of_radiance/io/backends/minimp4.h (the only native code of the reference), exactly as
 * io/mp4.nim does: MP4E_open(sequential 0, fragmentation 0) + mp4_h26x_write_init (mp4.nim:139-160), then every
 * NAL unit of the .264 file through mp4_h26x_write_nal with 90000 div fps ticks (get_nal_size / write_mp4,
 * mp4.nim:66-96), then MP4E_close + mp4_h26x_write_close (mp4.nim:109-113).
 *
 * TEST INFRASTRUCTURE ONLY.  Built by oracle/Makefile into oracle/_ref/ref_mp4mux when /root/reference exists; the
 * header is compiled where it lies and never copied into this repository.
 *   usage: ref_mp4mux <src.264> <dst.mp4> <width> <height> [fps]
 */
#define MINIMP4_IMPLEMENTATION
#include "minimp4.h" /* -I/root/reference/trace_of_radiance/io/backends */

static int write_cb(int64_t offset, const void* buffer, size_t size, void* token) { /* mp4.nim:116-126 */
  FILE* f = (FILE*)token;
  fseek(f, (long)offset, SEEK_SET);
  return fwrite(buffer, 1, size, f) != size;
}

static size_t nal_size(const unsigned char* buf, size_t size) { /* mp4.nim:66-74 */
  size_t pos = 3;
  while (size - pos > 3) {
    if (buf[pos] == 0 && buf[pos + 1] == 0 && buf[pos + 2] == 1) return pos;
    if (buf[pos] == 0 && buf[pos + 1] == 0 && buf[pos + 2] == 0 && buf[pos + 3] == 1) return pos;
    ++pos;
  }
  return size;
}

int main(int argc, char** argv) {
  if (argc < 5) {
    fprintf(stderr, "usage: %s <src.264> <dst.mp4> <width> <height> [fps]\n", argv[0]);
    return 2;
  }
  int width = atoi(argv[3]), height = atoi(argv[4]), fps = argc > 5 ? atoi(argv[5]) : 30;
  FILE* in = fopen(argv[1], "rb");
  if (!in) return 3;
  fseek(in, 0, SEEK_END);
  long n = ftell(in);
  fseek(in, 0, SEEK_SET);
  unsigned char* data = (unsigned char*)malloc((size_t)n);
  if (fread(data, 1, (size_t)n, in) != (size_t)n) return 3;
  fclose(in);
  FILE* out = fopen(argv[2], "wb");
  if (!out) return 4;
  MP4E_mux_t* mux = MP4E_open(0, 0, out, write_cb);
  mp4_h26x_writer_t wr;
  if (mp4_h26x_write_init(&wr, mux, width, height, 0) != MP4E_STATUS_OK) return 5;
  unsigned char* p = data;
  size_t left = (size_t)n;
  while (left > 0) { /* mp4.nim:81-96 */
    size_t sz = nal_size(p, left);
    if (sz < 4) {
      p += 1;
      left -= 1;
      continue;
    }
    if (mp4_h26x_write_nal(&wr, p, (int)sz, (unsigned)(90000 / fps)) != MP4E_STATUS_OK) return 6;
    p += sz;
    left -= sz;
  }
  MP4E_close(mux);
  mp4_h26x_write_close(&wr);
  fclose(out);
  free(data);
  return 0;
}

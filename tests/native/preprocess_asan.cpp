// Memory-safety run of the device code of csrc/preprocess.cu on the CPU: the emulation (preprocess_emu.cpp) driven over
// exact-size heap buffers under AddressSanitizer + UBSan, so that any access outside the source images, the workspace,
// the staged shared-memory rows or the output — and any signed overflow of the int32 accumulators — aborts.
// Built and run by tests/test_preprocess_cpu.py::test_device_code_is_memory_safe.
#include <stdio.h>

#include "preprocess_emu.cpp"

static int run(const int c[9]) {
  const int n = c[0], h = c[1], w = c[2], oh = c[7], ow = c[8];
  const size_t src_bytes = static_cast<size_t>(n) * h * w * 3;
  uint8_t* src = static_cast<uint8_t*>(malloc(src_bytes));
  for (size_t i = 0; i < src_bytes; ++i) src[i] = (i / 7) % 2 ? 255 : 0;  // saturated stripes: widest accumulators
  float* out = static_cast<float*>(malloc(static_cast<size_t>(n) * 3 * oh * ow * sizeof(float)));
  const float half[3] = {0.5f, 0.5f, 0.5f};
  const int rc = emu_image_preprocess(src, n, h, w, c[3], c[4], c[5], c[6], oh, ow, half, half, out, nullptr);
  uint8_t* u8 = static_cast<uint8_t*>(malloc(static_cast<size_t>(n) * oh * ow * 3));
  emu_image_postprocess(out, n, oh, ow, half, half, u8);
  free(src), free(out), free(u8);
  return rc;
}

int main() {
  // n, in_h, in_w, res_h, res_w, crop_top, crop_left, out_h, out_w
  const int cases[][9] = {{2, 517, 389, 340, 256, 42, 0, 256, 256},  {1, 300, 451, 512, 769, 0, 128, 512, 512},
                          {3, 3, 301, 5, 299, 0, 0, 5, 299},         {1, 40, 2600, 40, 3, 0, 0, 40, 3},
                          {2, 97, 131, 64, 86, 0, 11, 64, 64},       {1, 131, 97, 86, 64, 11, 0, 64, 64},
                          {1, 64, 100, 64, 50, 0, 0, 64, 50},        {1, 100, 64, 50, 64, 0, 0, 50, 64},
                          {1, 33, 33, 7, 5, 0, 0, 7, 5},             {2, 64, 640, 64, 640, 0, 288, 64, 64},
                          {1, 1, 1, 5, 5, 0, 0, 5, 5},               {1, 5, 5, 1, 1, 0, 0, 1, 1},
                          {1, 7, 1000, 7, 333, 0, 100, 7, 200},      {1, 256, 256, 256, 256, 0, 0, 256, 256}};
  for (const auto& c : cases) {
    const int rc = run(c);
    if (rc != 0) {
      printf("case in %dx%d failed: %d\n", c[1], c[2], rc);
      return 1;
    }
  }
  printf("asan ok\n");
  return 0;
}

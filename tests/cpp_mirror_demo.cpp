// TEST-ONLY: exercises the C++ mirror of the reference API (aprilgrid-rs_b200/cpp/aprilgrid_b200.hpp)
// end to end on a GPU: reads a raw 8-bit gray frame (width height on the command line), runs
// TagDetector::detect, detect_batch on three copies and MultiTagDetector::detect_batch, and prints
// "id x0 y0 ... x3 y3" lines (sorted by id) followed by the batch consistency verdict.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "aprilgrid_b200.hpp"

int main(int argc, char** argv) {
  if (argc < 4) return 2;
  const int w = atoi(argv[2]), h = atoi(argv[3]);
  std::vector<unsigned char> img((size_t)w * h);
  FILE* f = fopen(argv[1], "rb");
  if (!f || fread(img.data(), 1, img.size(), f) != img.size()) return 3;
  fclose(f);
  try {
    aprilgrid::TagDetector det(aprilgrid::tag_family_from_str("t36h11"));
    const aprilgrid::TagMap tags = det.detect({img.data(), w, h, 0, AG_L8});
    std::vector<uint32_t> ids;
    for (auto& kv : tags) ids.push_back(kv.first);
    std::sort(ids.begin(), ids.end());
    for (uint32_t id : ids) {
      printf("%u", id);
      for (auto& c : tags.at(id)) printf(" %.9g %.9g", c.first, c.second);
      printf("\n");
    }
    std::vector<unsigned char> three;
    for (int i = 0; i < 3; ++i) three.insert(three.end(), img.begin(), img.end());
    const auto batch = det.detect_batch(three.data(), img.size(), 3, w, h, 0, AG_L8);
    aprilgrid::MultiTagDetector multi(aprilgrid::TagFamily::T36H11);
    const auto mb = multi.detect_batch(three.data(), img.size(), 3, w, h, 0, AG_L8);
    bool same = batch.size() == 3 && mb.size() == 3;
    for (int i = 0; same && i < 3; ++i) same = batch[i] == tags && mb[i] == tags;
    printf("batch %s saddles %zu devices %d\n", same ? "same" : "DIFFERENT", det.refined_saddle_points({img.data(), w, h, 0, AG_L8}).size(),
           multi.device_count());
  } catch (const std::exception& e) {
    fprintf(stderr, "%s\n", e.what());
    return 1;
  }
  return 0;
}

/* TEST INFRASTRUCTURE — one-off extractor of dlib's OWN golden vectors for FHOG:
 * dlib/test/fhog.cpp:156-213 embeds (base64 + compress_stream) a face image and the expected
 * array2d<matrix<float,31,1>> outputs for two cell sizes, asserted there to 1e-6.  This tool
 * includes that test file in place from /root/reference (nothing is copied), decodes the blobs with
 * dlib's own routines and writes them as raw little-endian arrays; tests/golden/make_golden_dlib.py
 * turns them into tests/golden/fhog_dlib_face.npz.  Build + run: see that script. */
#include <sstream>
#include <string>
#include <cstdlib>
#include <ctime>
#include <fstream>
#include <dlib/image_transforms.h>
#include <dlib/image_io.h>
#include <dlib/compress_stream.h>
#include <dlib/base64.h>
#include "tester.h"
#define private public
#define protected public
#include "fhog.cpp"          /* resolved through -I<ref>/dlib/test */
#undef private
#undef protected
#include <fstream>

int main(int argc, char **argv) {
  using namespace dlib;
  using namespace std;
  fhog_tester &t = a;      /* the file-scope test instance */
  array2d<rgb_pixel> img;
  istringstream sin(t.get_decoded_string_face_dng());
  load_dng(img, sin);
  sin.clear(); sin.str(t.get_decoded_string_fhog_feats());
  int sbin1, sbin2;
  array2d<matrix<float, 31, 1> > v1, v2;
  deserialize(sbin1, sin); deserialize(v1, sin); deserialize(sbin2, sin); deserialize(v2, sin);
  ofstream f(argc > 1 ? argv[1] : "fhog_dlib_face.bin", ios::binary);
  auto wi = [&](int v) { f.write((const char *)&v, 4); };
  wi((int)img.nr()); wi((int)img.nc());
  for (long r = 0; r < img.nr(); r++) for (long c = 0; c < img.nc(); c++) { f.put(img[r][c].red); f.put(img[r][c].green); f.put(img[r][c].blue); }
  for (int k = 0; k < 2; k++) {
    array2d<matrix<float, 31, 1> > &v = k ? v2 : v1;
    wi(k ? sbin2 : sbin1); wi((int)v.nr()); wi((int)v.nc());
    for (long r = 0; r < v.nr(); r++) for (long c = 0; c < v.nc(); c++) for (int o = 0; o < 31; o++) { float x = v[r][c](o); f.write((const char *)&x, 4); }
  }
  return 0;
}

/* TEST INFRASTRUCTURE — the one C++ piece of the oracle: get_surf_points orders key points with
 *   std::sort(points.rbegin(), points.rend());            (dlib/image_keypoint/surf.h:268)
 * using interest_point::operator< on the score (hessian_pyramid.h:32).  std::sort is not stable, so
 * exactly tied scores come out in an order that only the same algorithm reproduces; the permutation
 * depends on the comparison results alone, hence sorting this mirror struct with libstdc++'s
 * std::sort gives the reference's order bit for bit. */
#include <algorithm>
#include <vector>

struct ip_mirror {
  double x, y, scale, score, lap;
  bool operator<(const ip_mirror &p) const { return score < p.score; }
};

extern "C" void orc_sort_points_like_reference(ip_mirror *p, int n) {
  std::vector<ip_mirror> v(p, p + n);
  std::sort(v.rbegin(), v.rend());
  std::copy(v.begin(), v.end(), p);
}

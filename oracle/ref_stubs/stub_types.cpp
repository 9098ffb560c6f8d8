// TEST INFRASTRUCTURE (oracle/_ref build only; never linked into the product).
// Replaces /root/reference/src/core/types.cpp, which needs <proj.h> (PROJ is not
// in this image).  Only the PROJ-free pieces are supplied: data_type_size and
// BBox follow types.cpp:11-43 semantics (contains() inclusive on all four edges);
// CRS is metadata only on the ingest/finalize path, so its methods are trivial.
#include "pcr/core/types.h"
#include <algorithm>

namespace pcr {

size_t data_type_size(DataType dt) {
    static const size_t sz[] = {4, 8, 4, 4, 2, 2, 1};
    unsigned i = static_cast<unsigned>(dt);
    return i < 7 ? sz[i] : 0;
}

void BBox::expand(double x, double y) {
    if (x < min_x) min_x = x;
    if (y < min_y) min_y = y;
    if (x > max_x) max_x = x;
    if (y > max_y) max_y = y;
}

void BBox::expand(const BBox& o) {
    if (!o.valid()) return;
    expand(o.min_x, o.min_y);
    expand(o.max_x, o.max_y);
}

bool BBox::contains(double x, double y) const {
    return !(x < min_x) && !(x > max_x) && !(y < min_y) && !(y > max_y)
           && x == x && y == y;
}

bool CRS::is_projected() const  { return false; }
bool CRS::is_geographic() const { return false; }
CRS  CRS::from_epsg(int code)   { CRS c; c.epsg = code; return c; }
CRS  CRS::from_wkt(const std::string& w) { CRS c; c.wkt = w; return c; }
bool CRS::equivalent_to(const CRS& o) const {
    return epsg != 0 && epsg == o.epsg;
}

}  // namespace pcr

// TEST INFRASTRUCTURE (oracle/_ref build only).  Replaces
// /root/reference/src/io/grid_io.cpp (GDAL is not in this image): oracle runs
// leave PipelineConfig.output_path empty, so these are never reached.
#include "pcr/io/grid_io.h"

namespace pcr {

Status write_geotiff(const std::string&, const Grid&, const GridConfig&,
                     const GeoTiffOptions&) {
    return Status::error(StatusCode::NotImplemented, "oracle build: no GDAL");
}

Status read_geotiff_info(const std::string&, int&, int&, int&, CRS&, BBox&) {
    return Status::error(StatusCode::NotImplemented, "oracle build: no GDAL");
}

}  // namespace pcr

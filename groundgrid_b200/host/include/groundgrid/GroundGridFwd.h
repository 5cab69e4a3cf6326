#pragma once
#include <memory>
namespace groundgrid {
class GroundGrid;
typedef std::shared_ptr<GroundGrid> GroundGridPtr;
typedef std::shared_ptr<const GroundGrid> GroundGridConstPtr;
}  // namespace groundgrid

// Stand-in for <boost/shared_ptr.hpp> (QPBaseClass.h:16,71 only needs reset()
// and operator->).  Test infrastructure only.
#pragma once
#include <memory>
namespace boost { template <class T> using shared_ptr = std::shared_ptr<T>; }

/* TEST INFRASTRUCTURE ONLY - see oracle/_shim/armawrap/newmat.h. boost::shared_ptr -> std::shared_ptr. */
#ifndef FABBER_SHIM_BOOST_SHARED_PTR
#define FABBER_SHIM_BOOST_SHARED_PTR
#include <memory>
namespace boost
{
using std::shared_ptr;
}
#endif

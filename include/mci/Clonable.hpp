// Clonable — same ownership convention as the reference (include/mci/Clonable.hpp:28-41): objects passed to MCI by
// reference are cloned, objects passed as unique_ptr are moved in.
#ifndef MCIG_MCI_CLONABLE_HPP
#define MCIG_MCI_CLONABLE_HPP
#include <memory>
namespace mci
{
template <typename T>
struct Clonable
{
protected:
    virtual T * _clone() const = 0;

public:
    virtual ~Clonable() = default;
    std::unique_ptr<T> clone() const { return std::unique_ptr<T>(_clone()); }
};
} // namespace mci
#endif

// storage of the stub's globals (adapter/stubs/OpenFOAMStub.H)
#include "OpenFOAMStub.H"
namespace Foam
{
FatalErrorStub FatalError;
std::ostream& Info = std::cout;
}

"""B200-native FSP right-hand-side path of NumCME.jl behind the reference's API.

The directory name contains a dot, so load it with ``__graft_entry__.load_package()`` (registers the
package as ``numcme_jl_b200``).  Mutating Julia functions ``f!`` are spelled ``f_`` here.
"""
from ._lib import ArgumentError, NcmeError, SeparabilityError, LIB_PATH, load as load_library
from .device import Context, DeviceVector
from .cmemodel import (CmeModel, CmeModelWithSensitivity, Propensity, StandardTimeInvariantPropensity,
                       SeparableTimeVaryingPropensity, JointTimeVaryingPropensity, propensity, propensitygrad,
                       propensitygrad_timevarying, istimevarying, istimeseparable, get_parameters, get_stoich_matrix,
                       get_propensities, get_species_count, get_reaction_count, get_parameter_count,
                       get_propensity_gradients, get_gradient_sparsity_patterns)
from .statespace import (StateSpaceSparse, expand_, deleteat_, get_state_count, get_sink_count, get_states,
                         get_statedict, get_state_connectivity, get_sink_connectivity)
from .fspmatrix import FspMatrixSparse, matvec_, matvecadd_, matvec, get_rowcount, get_colcount
from .sensmatrix import ForwardSensFspMatrixSparse, sens_matvec_
from .fspvector import FspVectorSparse, FspOutputSparse, FspOutputSliceSparse, get_values, nnz
from .transientcme import (solve, AdaptiveFspSparse, RStepAdapter, SelectiveRStepAdapter, NativeRK45, NativeBDF, NativeBDFClassic,
                           NativeBDFFused, init_, adapt_)
from .forwardsenscme import (ForwardSensFspInitialConditionSparse, forwardsens_initial_condition, ForwardSensRStepAdapter,
                             get_probability, get_sensitivity,
                             AdaptiveForwardSensFspSparse, ForwardSensFspOutputSparse, ForwardSensFspOutputSliceSparse)
from .parallel import Comm, ShardedVector, shard_bounds
from . import workloads

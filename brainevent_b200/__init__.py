"""brainevent_b200 -- B200-native (sm_100a) event-driven synaptic propagation, a drop-in for
that one path of chaobrain/brainevent.  Public names follow brainevent/__init__.py."""
from . import config
from ._csr import (CSC, CSR, binary_csrmm, binary_csrmm_indexed, binary_csrmm_indexed_p, binary_csrmm_p,
                   binary_csrmm_p_call, binary_csrmv, binary_csrmv_auto, binary_csrmv_indexed, binary_csrmv_indexed_p,
                   binary_csrmv_p, binary_csrmv_p_call, csr_to_csc_index)
from ._dense import (binary_densemm, binary_densemm_p, binary_densemm_p_call, binary_densemv,
                     binary_densemv_p, binary_densemv_p_call)
from ._error import (BrainEventError, KernelError, KernelExecutionError, KernelLoadError,
                     KernelNotAvailableError, KernelRegistrationError, MathError,
                     UnsupportedOperationError)
from ._event import (BinaryArray, BitPackedBinary, CompactBinary, EventRepresentation,
                     binary_1d_array_index_p, binary_1d_array_index_p_call, binary_2d_array_index_p,
                     binary_2d_array_index_p_call, binary_2d_compact_only_p,
                     binary_2d_compact_only_p_call, bitpack)
from ._fcn import (FixedNumConn, FixedNumPerPost, FixedNumPerPre, binary_fcnmm, binary_fcnmm_p,
                   binary_fcnmm_p_call, binary_fcnmv, binary_fcnmv_p, binary_fcnmv_p_call)
from ._float import (csrmm, csrmm_p, csrmm_p_call, csrmv, csrmv_p, csrmv_p_call, fcnmm, fcnmm_p, fcnmm_p_call,
                     fcnmv, fcnmv_p, fcnmv_p_call, jitnmm, jitnmm_p, jitnmv, jitnmv_p, jitsmm, jitsmm_p, jitsmv,
                     jitsmv_p, jitumm, jitumm_p, jitumv, jitumv_p)
from ._jitc import (JITCMatrix, JITCNormalC, JITCNormalR, JITCScalarC, JITCScalarR, JITCUniformC,
                    JITCUniformR, binary_jitnmm, binary_jitnmm_p, binary_jitnmv, binary_jitnmv_p,
                    binary_jitsmm, binary_jitsmm_p, binary_jitsmv, binary_jitsmv_p, binary_jitumm,
                    binary_jitumm_p, binary_jitumv, binary_jitumv_p, jitc_row_connections, jitn_to_csr,
                    jits_to_csr, jitu_to_csr)
from ._encode import (binary_2d_csc_encode_p, binary_2d_csc_encode_p_call, binary_2d_csr_encode_p,
                      binary_2d_csr_encode_p_call, binary_2d_csr_fill_p_call, binary_2d_csr_row_count_p_call,
                      binary_2d_pair_stream_encode_p, binary_2d_pair_stream_encode_p_call,
                      binary_2d_row_sparse_encode_p, binary_2d_row_sparse_encode_p_call)
from ._plasticity import (update_csc_on_binary_post, update_csc_on_binary_pre, update_csr_on_binary_post,
                          update_csr_on_binary_pre, update_dense_on_binary_post, update_dense_on_binary_pre,
                          update_fixed_post_conn_on_binary_pre, update_fixed_pre_conn_on_binary_post)
from ._op import (KernelEntry, XLACustomKernel, get_all_primitive_names, get_primitives_by_tags,
                  get_registry)

__version__ = "0.1.0"

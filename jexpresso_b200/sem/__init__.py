"""Host-side SEM setup: the arrays Jexpresso's sem_setup/params_setup hand to rhs!."""
from .basis import build_basis, lgl_nodes_weights, lagrange_basis
from .mesh import BoxSpec, Mesh, structured_box, compute_xy_partition, part_subbox, element_sizes, effective_delta_l
from .setup import SEM, sem_setup, assemble_host, conformity4ncf_q_host
from .cases import rtb_initial_state

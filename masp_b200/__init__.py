"""masp_b200: Blackwell-native Groth16 proving path for the MASP Spend / Output /
Convert circuits, behind the masp_proofs proving surface.  See DESIGN.md."""
from . import circuits, sapling, synthetic  # noqa: F401
from .prover import (  # noqa: F401
    GROTH_PROOF_SIZE, LocalTxProver, Mb200Error, Parameters, ParameterError, ProvingAssignment, create_proof,
    create_proof_batch, create_random_proof, init, load_parameters, parse_parameters)
from .sapling import BatchingTxProver, SaplingError, SaplingProvingContext, TxProver  # noqa: F401,E402

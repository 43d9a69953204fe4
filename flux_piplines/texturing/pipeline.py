"""Drop-in module path of the reference's `flux_piplines/texturing/pipeline.py` (the reference's texturing and delight
copies are byte-identical): re-exports the B200-native `PBRFluxPipeline` from unitex_b200.flux_pipeline."""
from unitex_b200.flux_pipeline import (FlowMatchEulerSchedule, PBRFluxPipeline, PBRFluxPipelineOutput,  # noqa: F401
                                       calculate_shift)

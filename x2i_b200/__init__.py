"""x2i_b200 -- B200-native (sm_100a) implementation of the X2I hot path: the FLUX MMDiT denoise step, the alignment
projector and the attention-distillation loss, behind the reference's own Python interfaces.  See DESIGN.md."""
__version__ = "0.1.0"

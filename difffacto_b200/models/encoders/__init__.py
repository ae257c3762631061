from .part_encoders import PartAlignerTransformer, PartEncoderForTransformerDecoder, build_latent_flow  # noqa: F401
from .pointnet import PointNetV2  # noqa: F401

from .unet_3d import UNet3DConditionModel  # noqa: F401
from .mutual_self_attention import ReferenceAttentionControl  # noqa: F401
from .unet_2d_condition import UNet2DConditionModel  # noqa: F401
from .prologue import AudioProjection, VKpsGuider  # noqa: F401

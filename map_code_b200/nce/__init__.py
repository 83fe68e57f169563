from .index_linear import IndexLinear
from .nce_loss import NCELoss
from .alias_multinomial import AliasMultinomial

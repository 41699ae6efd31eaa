"""Hyper-parameters the model classes read (the reference's ``config/config.py`` dataclass, minus its side effects:
the reference's ``Config.__post_init__`` creates log / checkpoint directories in the cwd, :64-67).  Any object with these
attributes works, including the reference's own ``Config``."""
from dataclasses import dataclass


@dataclass
class Config:
    total_points: int = 120000          # config/config.py:19-20
    global_points: int = 30000
    time_embed_dim: int = 128           # :23-25
    feature_dim: int = 256
    global_feature_dim: int = 256
    num_timesteps: int = 1000           # :28-30
    beta_schedule: str = "cosine"
    noise_schedule_offset: float = 0.0008
    learning_rate: float = 1e-4         # :33-37
    weight_decay: float = 1e-4
    ema_decay: float = 0.999
    gradient_clip: float = 1.0
    cond_drop_prob: float = 0.1         # :40-41
    guidance_scale: float = 7.5
    use_amp: bool = True                # :51-52
    gradient_accumulation_steps: int = 3
    use_hierarchical: bool = True       # :60-62
    lambda_chamfer: float = 0.1

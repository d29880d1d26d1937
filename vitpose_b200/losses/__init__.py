from .mse_loss import JointsMSELoss

__all__ = ['JointsMSELoss']

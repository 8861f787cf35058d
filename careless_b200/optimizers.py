"""Adam hyper-parameter holder with tf_keras's argument names (careless/io/manager.py:494-501)."""


class Adam:
    def __init__(self, learning_rate=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-7, clipnorm=None, clipvalue=None,
                 global_clipnorm=None):
        if clipnorm is not None and global_clipnorm is not None:
            raise ValueError("Cannot accept both `clipnorm` and `global_clipnorm`")        # keras behaviour
        self.learning_rate, self.beta_1, self.beta_2, self.epsilon = learning_rate, beta_1, beta_2, epsilon
        self.clipnorm, self.clipvalue, self.global_clipnorm = clipnorm, clipvalue, global_clipnorm

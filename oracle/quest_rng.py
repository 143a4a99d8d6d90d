"""Restatement of the host RNG behind QuEST's measurements -- TEST INFRASTRUCTURE ONLY.

core/randomiser.cpp:48-82 keeps one std::mt19937_64 seeded through std::seed_seq(seeds); every measurement draws
exactly one std::uniform_real_distribution<double>(0,1) variate (randomiser.cpp:125-139) and returns `sample >
probOfZero`.  Reproducing that stream bit for bit makes measurement OUTCOMES part of the oracle (the reference's own
tests do not pin them).  Algorithms: ISO C++ [rand.util.seedseq], [rand.eng.mers] (mt19937_64 constants),
and libstdc++'s generate_canonical (one 64-bit draw, divided by 2^64, clamped below 1)."""

M32 = (1 << 32) - 1
M64 = (1 << 64) - 1


def seed_seq_generate(seeds, n):
    """std::seed_seq(seeds).generate(): n 32-bit words"""
    v = [s & M32 for s in seeds]
    s = len(v)
    b = [0x8B8B8B8B] * n
    t = 11 if n >= 623 else 7 if n >= 68 else 5 if n >= 39 else 3 if n >= 7 else (n - 1) // 2
    p = (n - t) // 2
    q = p + t
    m = max(s + 1, n)

    def T(x):
        return x ^ (x >> 27)

    for k in range(m):
        r1 = (1664525 * T(b[k % n] ^ b[(k + p) % n] ^ b[(k - 1) % n])) & M32
        if k == 0:
            r2 = (r1 + s) & M32
        elif k <= s:
            r2 = (r1 + (k % n) + v[k - 1]) & M32
        else:
            r2 = (r1 + (k % n)) & M32
        b[(k + p) % n] = (b[(k + p) % n] + r1) & M32
        b[(k + q) % n] = (b[(k + q) % n] + r2) & M32
        b[k % n] = r2
    for k in range(m, m + n):
        r3 = (1566083941 * T((b[k % n] + b[(k + p) % n] + b[(k - 1) % n]) & M32)) & M32
        r4 = (r3 - (k % n)) & M32
        b[(k + p) % n] ^= r3
        b[(k + q) % n] ^= r4
        b[k % n] = r4
    return b


class MT19937_64:
    N, M, R = 312, 156, 31
    A = 0xB5026F5AA96619E9
    U, D = 29, 0x5555555555555555
    S, B = 17, 0x71D67FFFEDA60000
    T, C = 37, 0xFFF7EEE000000000
    L = 43

    def __init__(self, seeds):
        """mersenne_twister_engine::seed(seed_seq&): two 32-bit words per state word, low word first"""
        a = seed_seq_generate(seeds, self.N * 2)
        self.x = [(a[2 * i] | (a[2 * i + 1] << 32)) & M64 for i in range(self.N)]
        upper = (M64 >> self.R) << self.R
        if (self.x[0] & upper) == 0 and all(w == 0 for w in self.x[1:]):
            self.x[0] = 1 << 63
        self.i = self.N

    def _twist(self):
        upper = (M64 >> self.R) << self.R
        lower = M64 ^ upper
        x, N, M = self.x, self.N, self.M
        for k in range(N):
            y = (x[k] & upper) | (x[(k + 1) % N] & lower)
            x[k] = x[(k + M) % N] ^ (y >> 1) ^ (self.A if (y & 1) else 0)
        self.i = 0

    def next_u64(self):
        if self.i >= self.N:
            self._twist()
        z = self.x[self.i]
        self.i += 1
        z ^= (z >> self.U) & self.D
        z ^= (z << self.S) & self.B & M64
        z ^= (z << self.T) & self.C & M64
        z ^= z >> self.L
        return z & M64

    def uniform01(self):
        """std::uniform_real_distribution<double>(0,1): generate_canonical<double,53> = double(x) / 2^64, < 1"""
        r = float(self.next_u64()) / 18446744073709551616.0
        return r if r < 1.0 else 1.0 - 2.0 ** -53

    def single_qubit_outcome(self, prob_of_zero):
        """rand_getRandomSingleQubitOutcome (randomiser.cpp:125-139)"""
        return int(self.uniform01() > prob_of_zero)

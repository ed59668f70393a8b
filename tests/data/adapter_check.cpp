#include "B200ModelRunner.h"
int main() { return 0; }

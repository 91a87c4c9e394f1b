package io.github.panghy.vectorsearch.gpu;

import static java.lang.foreign.ValueLayout.ADDRESS;
import static java.lang.foreign.ValueLayout.JAVA_INT;
import static java.lang.foreign.ValueLayout.JAVA_LONG;

import java.lang.foreign.Arena;
import java.lang.foreign.FunctionDescriptor;
import java.lang.foreign.Linker;
import java.lang.foreign.MemoryLayout;
import java.lang.foreign.MemorySegment;
import java.lang.foreign.SymbolLookup;
import java.lang.invoke.MethodHandle;
import java.util.concurrent.ConcurrentHashMap;

/**
 * Panama FFM binding of libvsgpu (include/vsgpu.h), one downcall handle per C entry point.
 *
 * <p>Needs JDK 22+ (java.lang.foreign is final there; the reference builds with JDK 21, where the same code compiles
 * with {@code --enable-preview}). The library is found through {@code -Dvsgpu.lib=/path/libvsgpu.so} or
 * {@code java.library.path}. Every C function returns an int status: {@link #check(int)} maps VS_EINVAL to
 * IllegalArgumentException (what PqTrainer.train throws, J/pq/PqTrainer.java:29-34), VS_EEMPTY to
 * IndexOutOfBoundsException (data.get(0) on an empty list, :49), anything else to IllegalStateException.
 */
public final class VsGpu {
  private VsGpu() {}

  public static final int VS_OK = 0, VS_EINVAL = -1, VS_ENOMEM = -2, VS_ECUDA = -3, VS_EHANDLE = -4, VS_ESTATE = -5, VS_EEMPTY = -6;
  public static final int METRIC_L2 = 0, METRIC_COSINE = 1;

  private static final Linker LINKER = Linker.nativeLinker();
  private static final SymbolLookup LIB = lookup();
  private static final ConcurrentHashMap<String, MethodHandle> HANDLES = new ConcurrentHashMap<>();

  private static SymbolLookup lookup() {
    String path = System.getProperty("vsgpu.lib");
    if (path != null) return SymbolLookup.libraryLookup(java.nio.file.Path.of(path), Arena.global());
    System.loadLibrary("vsgpu");
    return SymbolLookup.loaderLookup();
  }

  /** Downcall handle for {@code name}; arguments are described once, by the first caller. */
  public static MethodHandle fn(String name, MemoryLayout ret, MemoryLayout... args) {
    return HANDLES.computeIfAbsent(name, n -> LINKER.downcallHandle(
        LIB.find(n).orElseThrow(() -> new UnsatisfiedLinkError("libvsgpu does not export " + n)),
        ret == null ? FunctionDescriptor.ofVoid(args) : FunctionDescriptor.of(ret, args)));
  }

  /** int f(args...) with every argument an address, an int or a long, as given. */
  public static int call(String name, MemoryLayout[] layout, Object... args) {
    try {
      return (int) fn(name, JAVA_INT, layout).invokeWithArguments(args);
    } catch (RuntimeException | Error e) {
      throw e;
    } catch (Throwable t) {
      throw new IllegalStateException(t);
    }
  }

  public static String lastError() {
    try {
      MemorySegment p = (MemorySegment) fn("vs_last_error", ADDRESS).invokeExact();
      return p.reinterpret(4096).getString(0);
    } catch (Throwable t) {
      return "(vs_last_error failed: " + t + ")";
    }
  }

  public static void check(int rc) {
    if (rc == VS_OK) return;
    String msg = lastError();
    switch (rc) {
      case VS_EINVAL -> throw new IllegalArgumentException(msg);
      case VS_EEMPTY -> throw new IndexOutOfBoundsException(msg);
      case VS_ENOMEM -> throw new OutOfMemoryError(msg);
      default -> throw new IllegalStateException("libvsgpu error " + rc + ": " + msg);
    }
  }

  // layouts used over and over
  static final MemoryLayout[] PAIR = {ADDRESS, ADDRESS, JAVA_INT, ADDRESS};

  static MemoryLayout[] sig(MemoryLayout... l) {
    return l;
  }

  static final MemoryLayout A = ADDRESS, I = JAVA_INT, J = JAVA_LONG;
}

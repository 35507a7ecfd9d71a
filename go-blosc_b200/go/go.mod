// Same module path as the reference (/root/reference/go.mod:1): an application switches with a `replace`
// directive or by vendoring this directory, its import lines stay as they are.
module github.com/mrjoshuak/go-blosc

go 1.23
